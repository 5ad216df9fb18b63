// dd_tracker.cu -- sm_100a kernels + C ABI of the batched DeepSORT tick.
//
// Launch shapes (S streams, T = max_tracks, D = max_dets):
//   k_prep         one warp per (stream, detection)       4 warps / CTA
//   k_predict      one warp per (stream, track index)     4 warps / CTA
//   k_gate         one warp per (stream, track index)     4 warps / CTA
//   k_cosine       one warp per (stream, track index)     4 warps / CTA   <- the HBM-bound kernel
//   k_match        one warp per stream                    1 warp  / CTA, dynamic shared memory
//   k_apply        one warp per (stream, detection)       4 warps / CTA
//   k_countline    one warp per stream                    4 warps / CTA
#include <cuda_runtime.h>
#include "dd_tracker_bodies.cuh"
#include "dd_tma.cuh"

#define DD_WARPS 4

#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

// Small per-item kernels: DD_SUB lanes per item, 32 / DD_SUB items per warp (SubG).
#define DD_SUB 8
#define DD_ITEMS_PER_CTA (DD_WARPS * 32 / DD_SUB)

__global__ void __launch_bounds__(DD_WARPS * 32)
k_prep(const DDView V, const double* __restrict__ det_tlwh, const float* __restrict__ det_feat,
       const int* __restrict__ det_count) {
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (w >= V.S * V.D) return;
    SubG<DD_SUB> g;
    dd_prep_det(g, V, w / V.D, w % V.D, det_tlwh, det_feat, det_count);
}

__global__ void __launch_bounds__(DD_WARPS * 32) k_predict(const DDView V) {
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (w >= V.S * V.T) return;
    SubG<DD_SUB> g;
    dd_predict_track(g, V, w / V.T, w % V.T);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_gate(const DDView V, const int* __restrict__ det_count) {
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (w >= V.S * V.T) return;
    SubG<DD_SUB> g;
    dd_gate_track(g, V, w / V.T, w % V.T, det_count);
}

__global__ void __launch_bounds__(DD_WARPS * 32, 7)
k_cosine(const DDView V, const int* __restrict__ det_count) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.T) return;
    WarpG g;
    DDDirectPass<WarpG> pass;
    dd_cosine_track(g, V, w / V.T, w % V.T, det_count, pass);
}

// ---- TMA-staged gallery pass ---------------------------------------------------------------------
// Each warp owns a ring of DD_STAGES shared-memory stages of DD_ROWS gallery rows (4 KB) with one
// mbarrier per stage.  Lane 0 issues 1-D bulk copies (cp.async.bulk global -> shared, completion on the
// stage's mbarrier); loads in flight live in shared memory instead of registers, so a warp keeps
// DD_STAGES x 4 KB outstanding at ~70 registers/thread and the SM holds several such warps.
#define DD_STAGES 4
#define DD_STAGE_BYTES (DD_ROWS * DD_FEAT_DIM * 4)

struct DDTmaPass {
    float4* ring;                 // [DD_STAGES][DD_ROWS][32] float4, this warp's
    unsigned long long* bars;     // [DD_STAGES]
    unsigned phase;               // bit s = parity the next wait on stage s expects

    template <int NC>
    __device__ __forceinline__ void run(const WarpG& g, const float4* gal4, int glen,
                                        const float4* const (&qp)[DD_CH], float (&best)[DD_CH]) {
        constexpr int N = DD_ROWS * NC;
        float4 q[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) q[c] = qp[c][g.lane];
        float acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = -3.0e38f;
        const int nchunk = (glen + DD_ROWS - 1) / DD_ROWS;
        const char* src = (const char*)gal4;
        if (g.lane == 0) {
            const int pre = nchunk < DD_STAGES ? nchunk : DD_STAGES;
            for (int c = 0; c < pre; ++c) {
                const int rows = min(DD_ROWS, glen - c * DD_ROWS);
                dd_mbar_expect_tx(bars + c, rows * 512);
                dd_bulk_g2s(ring + c * (DD_ROWS * 32), src + (size_t)c * DD_STAGE_BYTES, rows * 512, bars + c);
            }
        }
        int st = 0;
        for (int ch = 0; ch < nchunk; ++ch) {
            dd_mbar_wait(bars + st, (phase >> st) & 1u);
            phase ^= 1u << st;
            const int rows = min(DD_ROWS, glen - ch * DD_ROWS);
            const float4* stage = ring + st * (DD_ROWS * 32);
            float4 a[DD_ROWS];
#pragma unroll
            for (int r = 0; r < DD_ROWS; ++r) a[r] = stage[min(r, rows - 1) * 32 + g.lane];
            __syncwarp();                                   // every lane has read the stage
            const int nxt = ch + DD_STAGES;
            if (g.lane == 0 && nxt < nchunk) {
                const int nrows = min(DD_ROWS, glen - nxt * DD_ROWS);
                dd_mbar_expect_tx(bars + st, nrows * 512);
                dd_bulk_g2s(ring + st * (DD_ROWS * 32), src + (size_t)nxt * DD_STAGE_BYTES, nrows * 512, bars + st);
            }
            float v[N];
#pragma unroll
            for (int r = 0; r < DD_ROWS; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float p = dd_fmaf(a[r].x, q[c].x, 0.f);
                    p = dd_fmaf(a[r].y, q[c].y, p);
                    p = dd_fmaf(a[r].z, q[c].z, p);
                    p = dd_fmaf(a[r].w, q[c].w, p);
                    v[r * NC + c] = p;
                }
            dd_fold_max<NC, N>(g, v, acc);
            st = (st + 1 == DD_STAGES) ? 0 : st + 1;
        }
        float b[NC];
        dd_fold_finish<NC, N>(g, acc, b);
#pragma unroll
        for (int c = 0; c < NC; ++c) best[c] = b[c];
    }
};

__global__ void __launch_bounds__(DD_WARPS * 32)
k_cosine_tma(const DDView V, const int* __restrict__ det_count) {
    extern __shared__ __align__(128) char smem[];
    const int wi = threadIdx.x >> 5;
    const int w = blockIdx.x * DD_WARPS + wi;
    if (w >= V.S * V.T) return;
    WarpG g;
    DDTmaPass pass;
    pass.ring = (float4*)(smem + (size_t)wi * DD_STAGES * DD_STAGE_BYTES);
    pass.bars = (unsigned long long*)(smem + (size_t)DD_WARPS * DD_STAGES * DD_STAGE_BYTES) + wi * DD_STAGES;
    pass.phase = 0;
    if (g.lane == 0) {
        for (int i = 0; i < DD_STAGES; ++i) dd_mbar_init(pass.bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    dd_cosine_track(g, V, w / V.T, w % V.T, det_count, pass);
}

static int g_gate_impl = 0;     // 1 = TMA-staged (default), 0 = direct loads (A/B baseline)

__global__ void __launch_bounds__(32)
k_match(const DDView V, const double* __restrict__ det_tlwh, const int* __restrict__ det_count,
        int* out_det_track_id) {
    extern __shared__ __align__(128) char smem[];
    WarpG g;
    dd_match_stream(g, V, blockIdx.x, det_tlwh, det_count, out_det_track_id, smem);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_apply(const DDView V, const float* __restrict__ det_conf, const int* __restrict__ det_label) {
    __shared__ double scratch[DD_ITEMS_PER_CTA][64];
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (w >= V.S * V.D) return;
    SubG<DD_SUB> g;
    dd_apply_det(g, V, w / V.D, w % V.D, det_conf, det_label, scratch[threadIdx.x / DD_SUB]);
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
static inline int items_to_blocks(long long n) { return (int)((n + DD_ITEMS_PER_CTA - 1) / DD_ITEMS_PER_CTA); }

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
    k_predict<<<items_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(V);
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
    const size_t smem = dd_match_smem_bytes(V.T, V.D, V.tab_cap);
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) return DD_ERR_INVALID;
        if (cudaFuncSetAttribute(k_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return DD_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[0], st);
    k_prep<<<items_to_blocks((long long)V.S * V.D), DD_WARPS * 32, 0, st>>>(V, det_tlwh, det_feat, det_count);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[1], st);
    k_gate<<<items_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st>>>(V, det_count);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[2], st);
    if (g_gate_impl == 1) {
        const size_t gsm = (size_t)DD_WARPS * DD_STAGES * DD_STAGE_BYTES + DD_WARPS * DD_STAGES * 8;
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(k_cosine_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm) != cudaSuccess)
                return DD_ERR_CUDA;
            attr_set = true;
        }
        k_cosine_tma<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, gsm, st>>>(V, det_count);
    } else {
        k_cosine<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st>>>(V, det_count);
    }
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[3], st);
    k_match<<<V.S, 32, smem, st>>>(V, det_tlwh, det_count, out_det_track_id);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[4], st);
    k_apply<<<items_to_blocks((long long)V.S * V.D), DD_WARPS * 32, 0, st>>>(V, det_conf, det_label);
    DD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[5], st);
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
                               void* const* host_events6) {
    if (!host_events6) return DD_ERR_INVALID;
    cudaEvent_t ev[6];
    for (int i = 0; i < 6; ++i) ev[i] = (cudaEvent_t)host_events6[i];
    return dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count,
                          out_det_track_id, (cudaStream_t)stream, ev);
}

int dd_tuning_set(int32_t key, int32_t value) {
    if (key == 0 && (value == 0 || value == 1)) { g_gate_impl = value; return DD_OK; }
    return DD_ERR_INVALID;
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
    k_predict<<<items_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st>>>(V);
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
