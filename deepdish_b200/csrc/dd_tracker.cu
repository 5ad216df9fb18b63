// dd_tracker.cu -- sm_100a kernels + C ABI of the batched DeepSORT tick.
//
// Launch shapes (S streams, T = max_tracks, D = max_dets):
//   k_prep         one warp per (stream, detection)       4 warps / CTA
//   k_predict      one warp per (stream, track index)     4 warps / CTA
//   k_gate_cosine  one warp per (stream, track index)     4 warps / CTA   <- the HBM-bound kernel
//   k_match        one warp per stream                    1 warp  / CTA, dynamic shared memory
//   k_apply        one warp per (stream, detection)       4 warps / CTA
//   k_countline    one warp per stream                    4 warps / CTA
#include <cuda_runtime.h>
#include "dd_tracker_bodies.cuh"

#define DD_WARPS 4

#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

__global__ void __launch_bounds__(DD_WARPS * 32)
k_prep(const DDView V, const double* __restrict__ det_tlwh, const float* __restrict__ det_feat,
       const int* __restrict__ det_count) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.D) return;
    WarpG g;
    dd_prep_det(g, V, w / V.D, w % V.D, det_tlwh, det_feat, det_count);
}

__global__ void __launch_bounds__(DD_WARPS * 32) k_predict(const DDView V) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.T) return;
    WarpG g;
    dd_predict_track(g, V, w / V.T, w % V.T);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_gate_cosine(const DDView V, const int* __restrict__ det_count) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.T) return;
    WarpG g;
    dd_gate_cosine(g, V, w / V.T, w % V.T, det_count);
}

__global__ void __launch_bounds__(32)
k_match(const DDView V, const double* __restrict__ det_tlwh, const int* __restrict__ det_count,
        int* out_det_track_id) {
    extern __shared__ __align__(16) char smem[];
    WarpG g;
    dd_match_stream(g, V, blockIdx.x, det_tlwh, det_count, out_det_track_id, smem);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_apply(const DDView V, const float* __restrict__ det_conf, const int* __restrict__ det_label) {
    __shared__ double scratch[DD_WARPS][64];
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.D) return;
    WarpG g;
    dd_apply_det(g, V, w / V.D, w % V.D, det_conf, det_label, scratch[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_countline(const DDView V, const double* __restrict__ line, int line_per_stream) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S) return;
    WarpG g;
    dd_countline(g, V, w, line + (line_per_stream ? (size_t)w * 4 : 0));
}

// counts [S, C*4] -> out [C*4]; one CTA per output element, tree reduction over streams.
__global__ void __launch_bounds__(256)
k_count_reduce(const long long* __restrict__ counts, int S, int n, long long* __restrict__ out) {
    __shared__ long long sh[256];
    const int e = blockIdx.x;
    long long acc = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) acc += counts[(size_t)s * n + e];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[e] = sh[0];
}

__global__ void k_init(const DDView V) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < V.S) V.next_id[s] = 1;
}

__global__ void __launch_bounds__(256)
k_status(const int* __restrict__ err, int S, int* __restrict__ out) {
    int acc = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) acc |= err[s];
    acc = __reduce_or_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicOr(out, acc);
}

static inline int warps_to_blocks(long long n_warps) { return (int)((n_warps + DD_WARPS - 1) / DD_WARPS); }

extern "C" {

const char* dd_version(void) { return "deepdish_b200 0.1.0 (sm_100a)"; }

int dd_tracker_layout_query(const dd_tracker_config* cfg, dd_tracker_layout* out) {
    return dd_layout_compute(cfg, out);
}

int dd_tracker_init(void* state, const dd_tracker_config* cfg, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    dd_tracker_layout L;
    dd_layout_compute(cfg, &L);
    cudaStream_t st = (cudaStream_t)stream;
    // everything except the gallery is zeroed; the gallery is only ever read below gal_len
    if (cudaMemsetAsync(state, 0, L.gal, st) != cudaSuccess) return DD_ERR_CUDA;
    if (cudaMemsetAsync((char*)state + L.lab_cnt, 0, L.total_bytes - L.lab_cnt, st) != cudaSuccess)
        return DD_ERR_CUDA;
    k_init<<<(V.S + 255) / 256, 256, 0, st>>>(V);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_tracker_predict(void* state, const dd_tracker_config* cfg, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    k_predict<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(V);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

static int dd_update_impl(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                          const float* det_conf, const int32_t* det_label, const float* det_feat,
                          const int32_t* det_count, int32_t* out_det_track_id, cudaStream_t st,
                          cudaEvent_t* ev) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!det_tlwh || !det_conf || !det_label || !det_feat || !det_count) return DD_ERR_INVALID;
    const size_t smem = dd_match_smem_bytes(V.T, V.D);
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) return DD_ERR_INVALID;
        if (cudaFuncSetAttribute(k_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return DD_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[0], st);
    k_prep<<<warps_to_blocks((long long)V.S * V.D), DD_WARPS * 32, 0, st>>>(V, det_tlwh, det_feat, det_count);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[1], st);
    k_gate_cosine<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st>>>(V, det_count);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[2], st);
    k_match<<<V.S, 32, smem, st>>>(V, det_tlwh, det_count, out_det_track_id);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[3], st);
    k_apply<<<warps_to_blocks((long long)V.S * V.D), DD_WARPS * 32, 0, st>>>(V, det_conf, det_label);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[4], st);
    return DD_OK;
}

int dd_tracker_update(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                      const float* det_conf, const int32_t* det_label, const float* det_feat,
                      const int32_t* det_count, int32_t* out_det_track_id, void* stream) {
    return dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count,
                          out_det_track_id, (cudaStream_t)stream, nullptr);
}

int dd_tracker_update_profiled(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                               const float* det_conf, const int32_t* det_label, const float* det_feat,
                               const int32_t* det_count, int32_t* out_det_track_id, void* stream,
                               void* const* host_events5) {
    if (!host_events5) return DD_ERR_INVALID;
    cudaEvent_t ev[5];
    for (int i = 0; i < 5; ++i) ev[i] = (cudaEvent_t)host_events5[i];
    return dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count,
                          out_det_track_id, (cudaStream_t)stream, ev);
}

int dd_event_create(void** host_out) {
    if (!host_out) return DD_ERR_INVALID;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return DD_ERR_CUDA;
    *host_out = (void*)e;
    return DD_OK;
}

int dd_event_destroy(void* ev) { return cudaEventDestroy((cudaEvent_t)ev) == cudaSuccess ? DD_OK : DD_ERR_CUDA; }

int dd_event_elapsed_ms(void* start, void* end, float* host_ms) {
    if (!host_ms) return DD_ERR_INVALID;
    return cudaEventElapsedTime(host_ms, (cudaEvent_t)start, (cudaEvent_t)end) == cudaSuccess ? DD_OK : DD_ERR_CUDA;
}

int dd_tracker_countline(void* state, const dd_tracker_config* cfg, const double* line,
                         int line_per_stream, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!line) return DD_ERR_INVALID;
    k_countline<<<warps_to_blocks(V.S), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(V, line, line_per_stream);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_tracker_tick(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                    const float* det_conf, const int32_t* det_label, const float* det_feat,
                    const int32_t* det_count, int32_t* out_det_track_id, const double* line,
                    int line_per_stream, int64_t* out_counts, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!line) return DD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    k_predict<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st>>>(V);
    DD_CHECK_LAUNCH();
    rc = dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count, out_det_track_id, st, nullptr);
    if (rc != DD_OK) return rc;
    k_countline<<<warps_to_blocks(V.S), DD_WARPS * 32, 0, st>>>(V, line, line_per_stream);
    DD_CHECK_LAUNCH();
    if (out_counts) {
        k_count_reduce<<<V.C * 4, 256, 0, st>>>(V.counts, V.S, V.C * 4, (long long*)out_counts);
        DD_CHECK_LAUNCH();
    }
    return DD_OK;
}

int dd_tracker_count_reduce(void* state, const dd_tracker_config* cfg, int64_t* out_counts, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!out_counts) return DD_ERR_INVALID;
    k_count_reduce<<<V.C * 4, 256, 0, (cudaStream_t)stream>>>(V.counts, V.S, V.C * 4, (long long*)out_counts);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_tracker_status(void* state, const dd_tracker_config* cfg, int32_t* host_flags, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!host_flags) return DD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int* d_out = nullptr;
    if (cudaMalloc(&d_out, sizeof(int)) != cudaSuccess) return DD_ERR_CUDA;
    cudaMemsetAsync(d_out, 0, sizeof(int), st);
    k_status<<<1, 256, 0, st>>>(V.err, V.S, d_out);
    cudaError_t e = cudaMemcpyAsync(host_flags, d_out, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_out);
    return e == cudaSuccess ? DD_OK : DD_ERR_CUDA;
}

}  // extern "C"
