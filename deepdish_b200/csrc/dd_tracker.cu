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
#ifndef DD_COSINE_MIN_CTAS
#define DD_COSINE_MIN_CTAS 5
#endif

#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

// Programmatic dependent launch (A/B knob dd_tuning_set(4, 1), off by default): the kernels of the tick are launched
// with programmatic stream serialisation (DDLaunch), so their CTAs may be scheduled while the previous kernel of the
// stream is still draining; the first thing each of them does is wait for that kernel's completion and memory flush,
// and only then let its own successor be scheduled.  Measured: 0.553 vs 0.558 ms per tick on one stream, nothing
// with two chunks, and the end-to-end path LOSES 20 % (1.44 vs 1.79 M stream-frames/s), so it stays off.  Without the
// launch attribute both instructions are no-ops.
__device__ __forceinline__ void dd_pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Small per-item kernels: DD_SUB lanes per item, 32 / DD_SUB items per warp (SubG).
#define DD_SUB 8
#define DD_ITEMS_PER_CTA (DD_WARPS * 32 / DD_SUB)

__global__ void __launch_bounds__(DD_WARPS * 32)
k_prep(const DDView V, const double* __restrict__ det_tlwh, const float* __restrict__ det_feat,
       const int* __restrict__ det_count) {
    dd_pdl_sync();
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (blockIdx.x == 0 && threadIdx.x == 0) {      // new tick: empty work list, claim cursor at 0
        V.work_ctl[0] = 0;
        V.work_ctl[32] = 0;
    }
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

// PREDICT = true: Tracker.predict of the same track index first (the fused tick): the 8 lanes that gate a track
// have just written its predicted mean / covariance, so the state is read back from L1 instead of HBM and one
// launch disappears.
// k_prep fed by a ragged blob (dd_unpack_detections' format): expands box / confidence / label / count into the
// padded arrays the later kernels read and normalises the feature straight from the blob, so the 512-byte feature
// rows are never copied to a padded staging buffer.
struct DDRagged {
    const unsigned char* blob;
    long long off_tlwh, off_conf, off_label, off_feat;
    double* det_tlwh;
    float* det_conf;
    int *det_label, *det_count;
};

__global__ void __launch_bounds__(DD_WARPS * 32)
k_prep_ragged(const DDView V, const DDRagged R) {
    dd_pdl_sync();
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (blockIdx.x == 0 && threadIdx.x == 0) {      // new tick: empty work list, claim cursor at 0
        V.work_ctl[0] = 0;
        V.work_ctl[32] = 0;
    }
    if (w >= V.S * V.D) return;
    SubG<DD_SUB> g;
    const int s = w / V.D, d = w - s * V.D;
    const int* offs = (const int*)R.blob;
    const int o0 = offs[s];
    const int n = offs[s + 1] - o0;
    if (d == 0 && g.lane == 0) R.det_count[s] = n;     // n > max_dets: the tick raises DD_FLAG_DET_OVERFLOW
    if (d >= n) return;
    const size_t src = (size_t)o0 + d, dst = (size_t)s * V.D + d;
    const double* box = (const double*)(R.blob + R.off_tlwh) + src * 4;
    if (g.lane < 4) R.det_tlwh[dst * 4 + g.lane] = box[g.lane];
    if (g.lane == 4) R.det_conf[dst] = ((const float*)(R.blob + R.off_conf))[src];
    if (g.lane == 5) R.det_label[dst] = ((const int*)(R.blob + R.off_label))[src];
    dd_prep_det_at(g, V, s, d, box, (const float*)(R.blob + R.off_feat) + src * DD_FEAT_DIM);
}

template <bool PREDICT>
__global__ void __launch_bounds__(DD_WARPS * 32)
k_gate(const DDView V, const int* __restrict__ det_count) {
    dd_pdl_sync();
    // gate, then append the track indices that have something to stream to the work list of the gallery
    // kernel: one atomicAdd per CTA (16 track indices), entries of a CTA stay in ascending order.
    __shared__ int s_has[DD_ITEMS_PER_CTA];
    __shared__ int s_base;
    const int item = threadIdx.x / DD_SUB;
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + item;
    SubG<DD_SUB> g;
    int has = 0;
    if (w < V.S * V.T) {
        if (PREDICT) {
            dd_predict_track(g, V, w / V.T, w % V.T);
            g.sync();
        }
        has = dd_gate_track(g, V, w / V.T, w % V.T, det_count) > 0 ? 1 : 0;
    }
    if (g.lane == 0) s_has[item] = has;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int h = threadIdx.x < DD_ITEMS_PER_CTA ? s_has[threadIdx.x] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, h != 0);
        if (threadIdx.x == 0) s_base = m ? atomicAdd(V.work_ctl, __popc(m)) : 0;
        if (threadIdx.x < DD_ITEMS_PER_CTA) s_has[threadIdx.x] = __popc(m & ((1u << threadIdx.x) - 1u));
    }
    __syncthreads();
    if (has && g.lane == 0) V.work[s_base + s_has[item]] = w;
}

__global__ void __launch_bounds__(DD_WARPS * 32, 7)
k_cosine(const DDView V, const int* __restrict__ det_count) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.T) return;
    WarpG g;
    DDDirectPass<WarpG> pass;
    dd_cosine_track(g, V, w / V.T, w % V.T, det_count, pass);
}

// Persistent form of the gallery kernel: a fixed grid (a few CTAs per SM, leaving registers and warp slots
// free so that the latency-bound kernels of OTHER stream chunks can run beside it), each warp claims track
// indices from the work list k_gate built until it is empty.  No idle warps (45 % of the track indices have
// nothing to stream), no CTA churn, a balanced tail; the pass itself is software-pipelined.
template <bool CS>
__global__ void __launch_bounds__(DD_WARPS * 32, DD_COSINE_MIN_CTAS)
k_cosine_work(const DDView V, const int* __restrict__ det_count) {
    dd_pdl_sync();
    WarpG g;
    const int n = V.work_ctl[0];
    DDPipelinedPass<CS> pass;
    for (;;) {
        int i = 0;
        if (g.lane == 0) i = atomicAdd(V.work_ctl + 32, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const int w = V.work[i];
        dd_cosine_track(g, V, w / V.T, w % V.T, det_count, pass);
    }
}


// ---- half-precision pre-pass + exact re-check ("k_cosine_h") ---------------------------------------------
// The gallery kernel is bound by HBM bytes, so it streams the round-to-nearest HALF copy of the gallery
// (galh, 256 B per row instead of 512) and uses it only to decide which rows can hold the exact maximum:
//   a(r, n) = tensor-core dot (mma.sync m16n8k16, f16 inputs, f32 accumulate) of half row r and half query n,
//   e(r, n) = the f32 value the exact pass computes (4 FMAs per lane + the 16-8-4-2-1 butterfly).
// |a - e| <= E := 1.1e-3 for unit vectors (2^-10 from rounding both operands to half, Cauchy-Schwarz; 1e-4 of
// slack for the tensor-core accumulation and 1e-5 for the f32 pass itself).  With m = max_r a(r, n) the row
// r* that maximises e satisfies a(r*, n) >= m - 2E, so the exact maximum is the maximum of e over the rows with
// a >= m - 2E -- typically one or two rows, read from the f32 gallery and evaluated with exactly the
// arithmetic of the exact pass.  The cost matrix is therefore bit-identical, whatever the data; only the number
// of re-checked rows (speed) depends on it.
// Fragment trick: a dot product does not care about the order of its terms, so each thread feeds the mma with
// the eight consecutive halves it loaded with one 16-byte load (lane = 4 g + t reads bytes 16 t + 64 j of row
// g and row g + 8, j = 0..3) and takes the B operand from the same bytes of query g: fully sectored loads, no
// shared-memory transposition.  Up to 8 gate-passing detections share one pass over the gallery.
#define DD_H_WINDOW 2.2e-3f

__device__ __forceinline__ void dd_mma_f16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                           unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct DDHalfSmem {          // per warp
    float* approx;           // [rows_pad][8]  a(r, n); overwritten with e(r, n) for the re-checked entries
    unsigned short* cand;    // [rows_pad * 8] re-check list, entry = row << 3 | n
    float* thr;              // [8]            m - 2E per query
    int* cj;                 // [8]            detection index of each query column
};
__host__ __device__ inline int dd_half_rows_pad(int B) { return (B + 15) & ~15; }
__host__ __device__ inline size_t dd_half_smem_per_warp(int B) {
    return (size_t)dd_half_rows_pad(B) * 8 * 6 + 64;
}

// one track index: all its gate-passing detections, 8 at a time
__device__ __forceinline__ void dd_cosine_track_half(const WarpG& g, const DDView& V, int s, int t,
                                                     const int* __restrict__ det_count, const DDHalfSmem& sm) {
    const int* desc = V.cdesc + ((size_t)s * V.T + t) * 2;
    if (desc[1] <= 0) return;
    const int d0 = desc[0];
    const size_t slot = (size_t)s * V.T + (d0 & 0xffff);
    const int glen = d0 >> 16;
    int nd = det_count[s];
    if (nd > V.D) nd = V.D;
    const int gq = g.lane >> 2, tq = g.lane & 3;
    const uint4* galh = (const uint4*)(V.galh + slot * (size_t)V.B * DD_FEAT_DIM);       // 16 uint4 per row
    const float4* gal4 = (const float4*)(V.gal + slot * (size_t)V.B * DD_FEAT_DIM);
    const int last = glen - 1;
    int base = 0;
    unsigned word = nd > 0 ? V.gate[slot * V.DW] : 0u;
    for (;;) {
        // ---- next group of <= 8 gate-passing detections
        int nq = 0;
        while (nq < 8) {
            if (!word) {
                base += 32;
                if (base >= nd) break;
                word = V.gate[slot * V.DW + (base >> 5)];
                continue;
            }
            if (g.lane == 0) sm.cj[nq] = base + dd_ctz(word);
            ++nq;
            word &= word - 1;
        }
        if (nq == 0) break;
        __syncwarp();
        const int myq = gq < nq ? sm.cj[gq] : 0;       // detection whose half row feeds column gq
        if (glen <= 0) {
            if (g.lane < nq) V.cost[slot * V.D + sm.cj[g.lane]] = dd_subf(1.0f, -3.0e38f);
            __syncwarp();
            continue;
        }
        uint4 qb[4];
        {
            const uint4* qh = (const uint4*)(V.det_feath + ((size_t)s * V.D + myq) * DD_FEAT_DIM);
#pragma unroll
            for (int j = 0; j < 4; ++j) qb[j] = gq < nq ? qh[tq + 4 * j] : make_uint4(0u, 0u, 0u, 0u);
        }
        // ---- stream the half gallery, 16 rows per step, two steps in flight
        const int nsteps = (glen + 15) >> 4;
        uint4 ga[2][4], gb[2][4];
#pragma unroll
        for (int st = 0; st < 2; ++st) {
            const uint4* r0 = galh + (size_t)dd_imin(st * 16 + gq, last) * 16 + tq;
            const uint4* r1 = galh + (size_t)dd_imin(st * 16 + gq + 8, last) * 16 + tq;
#pragma unroll
            for (int j = 0; j < 4; ++j) { ga[st][j] = r0[4 * j]; gb[st][j] = r1[4 * j]; }
        }
        float mx0 = -3.0e38f, mx1 = -3.0e38f;
        for (int step = 0; step < nsteps; step += 2) {
#pragma unroll
            for (int st = 0; st < 2; ++st) {
                const int cur = step + st;
                if (cur >= nsteps) break;
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dd_mma_f16(c, ga[st][j].x, gb[st][j].x, ga[st][j].y, gb[st][j].y, qb[j].x, qb[j].y);
                    dd_mma_f16(c, ga[st][j].z, gb[st][j].z, ga[st][j].w, gb[st][j].w, qb[j].z, qb[j].w);
                }
                const int nxt = cur + 2;
                if (nxt < nsteps) {
                    const uint4* r0 = galh + (size_t)dd_imin(nxt * 16 + gq, last) * 16 + tq;
                    const uint4* r1 = galh + (size_t)dd_imin(nxt * 16 + gq + 8, last) * 16 + tq;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { ga[st][j] = r0[4 * j]; gb[st][j] = r1[4 * j]; }
                }
                const int ra = cur * 16 + gq, rb = ra + 8;
                if (ra >= glen) { c[0] = -3.0e38f; c[1] = -3.0e38f; }      // padding rows never win
                if (rb >= glen) { c[2] = -3.0e38f; c[3] = -3.0e38f; }
                *(float2*)(sm.approx + ra * 8 + 2 * tq) = make_float2(c[0], c[1]);
                *(float2*)(sm.approx + rb * 8 + 2 * tq) = make_float2(c[2], c[3]);
                mx0 = fmaxf(mx0, fmaxf(c[0], c[2]));
                mx1 = fmaxf(mx1, fmaxf(c[1], c[3]));
            }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        }
        if (gq == 0) { sm.thr[2 * tq] = mx0 - DD_H_WINDOW; sm.thr[2 * tq + 1] = mx1 - DD_H_WINDOW; }
        __syncwarp();
        // ---- re-check list: every (row, query) whose approximate dot is within the window of the maximum
        int ncand = 0;
        const int total = nsteps * 16 * 8;
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + g.lane;
            const int n = i & 7;
            const bool p = n < nq && sm.approx[i] >= sm.thr[n];
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) sm.cand[ncand + __popc(m & ((1u << g.lane) - 1u))] = (unsigned short)i;
            ncand += __popc(m);
        }
        __syncwarp();
        // ---- exact values of the listed entries, 4 per round: the exact pass's arithmetic, bit for bit
        for (int c0 = 0; c0 < ncand; c0 += 4) {
            float v[4];
            float4 a[4], q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e = sm.cand[dd_imin(c0 + k, ncand - 1)];
                const int d = sm.cj[e & 7];
                a[k] = gal4[(size_t)(e >> 3) * (DD_FEAT_DIM / 4) + g.lane];
                q[k] = ((const float4*)(V.det_featn + ((size_t)s * V.D + d) * DD_FEAT_DIM))[g.lane];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float p = dd_fmaf(a[k].x, q[k].x, 0.f);
                p = dd_fmaf(a[k].y, q[k].y, p);
                p = dd_fmaf(a[k].z, q[k].z, p);
                p = dd_fmaf(a[k].w, q[k].w, p);
                v[k] = p;
            }
            int n = 4, o = 16;                          // transposing butterfly, as dd_fold_max
#pragma unroll
            for (; n > 1; n >>= 1, o >>= 1) {
                const bool up = (g.lane & o) != 0;
                const int half = n >> 1;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const float send = up ? v[i] : v[i + half];
                    const float keep = up ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
#pragma unroll
            for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
            const int k = g.lane >> 3;                  // lane L owns the total of entry c0 + (L >> 3)
            if ((g.lane & 7) == 0 && c0 + k < ncand) sm.approx[sm.cand[c0 + k]] = v[0];
        }
        __syncwarp();
        if (g.lane < nq) {                              // exact maximum per query over its re-checked rows
            float best = -3.0e38f;
            for (int i = 0; i < ncand; ++i) {
                const int e = sm.cand[i];
                if ((e & 7) == g.lane) best = fmaxf(best, sm.approx[e]);
            }
            V.cost[slot * V.D + sm.cj[g.lane]] = dd_subf(1.0f, best);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(DD_WARPS * 32, 4)
k_cosine_h(const DDView V, const int* __restrict__ det_count) {
    dd_pdl_sync();
    extern __shared__ __align__(16) char smem[];
    WarpG g;
    const int rows_pad = dd_half_rows_pad(V.B);
    char* mine = smem + (size_t)(threadIdx.x >> 5) * dd_half_smem_per_warp(V.B);
    DDHalfSmem sm;
    sm.approx = (float*)mine;
    sm.cand = (unsigned short*)(mine + (size_t)rows_pad * 8 * 4);
    sm.thr = (float*)(mine + (size_t)rows_pad * 8 * 6);
    sm.cj = (int*)(mine + (size_t)rows_pad * 8 * 6 + 32);
    const int n = V.work_ctl[0];
    for (;;) {
        int i = 0;
        if (g.lane == 0) i = atomicAdd(V.work_ctl + 32, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const int w = V.work[i];
        dd_cosine_track_half(g, V, w / V.T, w % V.T, det_count, sm);
    }
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

static int g_gate_impl = 3;     // gallery kernel: 2 = persistent work-list kernel (default), 0 = one warp per
                                // track index over the whole grid, 1 = TMA-staged ring (A/B baselines)
static int g_cosine_ctas_per_sm = 4;   // persistent grid = SMs x this
static int g_gallery_streaming = 0;    // 1: gallery loads are ld.global.cs (evict-first), 0: default policy
static int g_match_cta = -1;           // matching kernel: -1 = by problem size, 0 = one warp per stream, 1 = 4 warps
static int g_small_priority = 0;       // 1: launch the latency-bound kernels at the highest stream priority

static int dd_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// launch config carrying a per-kernel priority (cudaLaunchAttributePriority): the small latency-bound kernels
// of one stream chunk must not queue behind the not-yet-dispatched CTAs of another chunk's gallery kernel.
static int g_pdl = 0;                  // 1: programmatic dependent launch between the kernels of a tick (A/B knob)
struct DDLaunch {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    DDLaunch(unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool high) {
        static int lo = 0, hi = 0, have = 0;
        if (!have) { cudaDeviceGetStreamPriorityRange(&lo, &hi); have = 1; }
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        attr[0].id = cudaLaunchAttributePriority;
        attr[0].val.priority = (high && g_small_priority) ? hi : lo;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
    }
};

__global__ void __launch_bounds__(32)
k_match(const DDView V, const double* __restrict__ det_tlwh, const int* __restrict__ det_count,
        int* out_det_track_id) {
    dd_pdl_sync();
    extern __shared__ __align__(128) char smem[];
    WarpG g;
    dd_match_stream(g, V, blockIdx.x, det_tlwh, det_count, out_det_track_id, smem);
}

// Crowded scenes (C4: ~260 tracks x ~180 detections per stream): the same matching body run by 4 (or 8) warps of
// one CTA per stream, so every column scan, list compaction and staging loop is 4x (8x) wider.
template <int NW>
__global__ void __launch_bounds__(NW * 32)
k_match_cta(const DDView V, const double* __restrict__ det_tlwh, const int* __restrict__ det_count,
            int* out_det_track_id) {
    extern __shared__ __align__(128) char smem[];
    dd_pdl_sync();
    CtaG<NW> g(smem);
    dd_match_stream(g, V, blockIdx.x, det_tlwh, det_count, out_det_track_id, smem + 256);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_apply(const DDView V, const float* __restrict__ det_conf, const int* __restrict__ det_label) {
    dd_pdl_sync();
    __shared__ double scratch[DD_ITEMS_PER_CTA][64];
    const int w = blockIdx.x * DD_ITEMS_PER_CTA + threadIdx.x / DD_SUB;
    if (w >= V.S * V.D) return;
    SubG<DD_SUB> g;
    dd_apply_det(g, V, w / V.D, w % V.D, det_conf, det_label, scratch[threadIdx.x / DD_SUB]);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_countline(const DDView V, const double* __restrict__ line, int line_per_stream) {
    dd_pdl_sync();
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S) return;
    WarpG g;
    dd_countline(g, V, w, line + (line_per_stream ? (size_t)w * 4 : 0));
}

// counts [S, C*4] -> out [C*4]; one CTA per output element, tree reduction over streams.
__global__ void __launch_bounds__(256)
k_count_reduce(const long long* __restrict__ counts, int S, int n, long long* __restrict__ out) {
    dd_pdl_sync();
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

// Ragged host batch -> the padded arrays of the tick.  One warp per detection (512 B feature + box + conf + label).
__global__ void __launch_bounds__(DD_WARPS * 32)
k_unpack(const unsigned char* __restrict__ blob, int S, int dmax, long long off_tlwh, long long off_conf,
         long long off_label, long long off_feat, double* __restrict__ det_tlwh, float* __restrict__ det_conf,
         int* __restrict__ det_label, float* __restrict__ det_feat, int* __restrict__ det_count) {
    const int* offs = (const int*)blob;                           // [S + 1]
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= S * dmax) return;
    const int s = w / dmax, d = w - s * dmax;
    const int o0 = offs[s];
    const int n = offs[s + 1] - o0;
    if (d == 0 && lane == 0) det_count[s] = n;                    // n > dmax: the tick raises DD_FLAG_DET_OVERFLOW
    if (d >= n) return;
    const size_t src = (size_t)o0 + d, dst = (size_t)s * dmax + d;
    ((float4*)(det_feat + dst * DD_FEAT_DIM))[lane] = ((const float4*)(blob + off_feat) + src * (DD_FEAT_DIM / 4))[lane];
    if (lane < 4) det_tlwh[dst * 4 + lane] = ((const double*)(blob + off_tlwh))[src * 4 + lane];
    if (lane == 4) det_conf[dst] = ((const float*)(blob + off_conf))[src];
    if (lane == 5) det_label[dst] = ((const int*)(blob + off_label))[src];
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
    // everything except the gallery and its half copy (adjacent in the blob) is zeroed; both are only ever read
    // below gal_len
    if (L.galh != dd_align256(L.gal + 4ull * V.S * V.T * V.B * DD_FEAT_DIM) || L.lab_cnt < L.galh) return DD_ERR_INVALID;
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
                          cudaEvent_t* ev, cudaEvent_t gallery_wait = nullptr, cudaEvent_t gallery_done = nullptr,
                          bool with_predict = false, const DDRagged* ragged = nullptr) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!det_tlwh || !det_conf || !det_label || (!det_feat && !ragged) || !det_count) return DD_ERR_INVALID;
    const size_t smem = dd_match_smem_bytes(V.T, V.D, V.tab_cap);
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) return DD_ERR_INVALID;
        if (cudaFuncSetAttribute(k_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return DD_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[0], st);
    {
        DDLaunch L(items_to_blocks((long long)V.S * V.D), DD_WARPS * 32, 0, st, true);
        const cudaError_t le = ragged ? cudaLaunchKernelEx(&L.cfg, k_prep_ragged, V, *ragged)
                                      : cudaLaunchKernelEx(&L.cfg, k_prep, V, det_tlwh, det_feat, det_count);
        if (le != cudaSuccess) return DD_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[1], st);
    {
        DDLaunch L(items_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st, true);
        const cudaError_t le = with_predict ? cudaLaunchKernelEx(&L.cfg, k_gate<true>, V, det_count)
                                            : cudaLaunchKernelEx(&L.cfg, k_gate<false>, V, det_count);
        if (le != cudaSuccess) return DD_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[2], st);
    // stream chunks take turns on the HBM-bound gallery kernel: one at a time at full bandwidth, while the
    // latency-bound kernels of the other chunks run beside it in the SM resources its fixed grid leaves free
    if (gallery_wait && cudaStreamWaitEvent(st, gallery_wait, 0) != cudaSuccess) return DD_ERR_CUDA;
    if (g_gate_impl == 1) {
        const size_t gsm = (size_t)DD_WARPS * DD_STAGES * DD_STAGE_BYTES + DD_WARPS * DD_STAGES * 8;
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(k_cosine_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm) != cudaSuccess)
                return DD_ERR_CUDA;
            attr_set = true;
        }
        k_cosine_tma<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, gsm, st>>>(V, det_count);
    } else if (g_gate_impl == 3 && dd_half_smem_per_warp(V.B) * DD_WARPS <= 100 * 1024) {
        long long grid = (long long)dd_sm_count() * g_cosine_ctas_per_sm;
        const long long need = warps_to_blocks((long long)V.S * V.T);
        if (grid > need) grid = need;
        const size_t hsm = dd_half_smem_per_warp(V.B) * DD_WARPS;
        if (hsm > 48 * 1024 &&
            cudaFuncSetAttribute(k_cosine_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm) != cudaSuccess)
            return DD_ERR_CUDA;
        DDLaunch L((unsigned)grid, DD_WARPS * 32, hsm, st, false);
        if (cudaLaunchKernelEx(&L.cfg, k_cosine_h, V, det_count) != cudaSuccess) return DD_ERR_CUDA;
    } else if (g_gate_impl == 0) {
        k_cosine<<<warps_to_blocks((long long)V.S * V.T), DD_WARPS * 32, 0, st>>>(V, det_count);
    } else {
        long long grid = (long long)dd_sm_count() * g_cosine_ctas_per_sm;
        const long long need = warps_to_blocks((long long)V.S * V.T);
        if (grid > need) grid = need;
        DDLaunch L((unsigned)grid, DD_WARPS * 32, 0, st, false);
        const cudaError_t le = g_gallery_streaming ? cudaLaunchKernelEx(&L.cfg, k_cosine_work<true>, V, det_count)
                                                   : cudaLaunchKernelEx(&L.cfg, k_cosine_work<false>, V, det_count);
        if (le != cudaSuccess) return DD_ERR_CUDA;
    }
    DD_CHECK_LAUNCH();
    if (gallery_done && cudaEventRecord(gallery_done, st) != cudaSuccess) return DD_ERR_CUDA;
    if (ev) cudaEventRecord(ev[3], st);
    const int wide = g_match_cta >= 0 ? g_match_cta : ((V.T > 160 || V.D > 160) ? 1 : 0);
    if (wide && smem + 256 <= 227 * 1024) {
        const size_t wsm = smem + 256;
        if (wsm > 48 * 1024 &&
            (cudaFuncSetAttribute(k_match_cta<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm) != cudaSuccess ||
             cudaFuncSetAttribute(k_match_cta<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm) != cudaSuccess))
            return DD_ERR_CUDA;
        DDLaunch L(V.S, (wide == 2 ? 8 : 4) * 32, wsm, st, true);
        const cudaError_t le = wide == 2 ? cudaLaunchKernelEx(&L.cfg, k_match_cta<8>, V, det_tlwh, det_count, out_det_track_id)
                                         : cudaLaunchKernelEx(&L.cfg, k_match_cta<4>, V, det_tlwh, det_count, out_det_track_id);
        if (le != cudaSuccess) return DD_ERR_CUDA;
    } else {
        DDLaunch L(V.S, 32, smem, st, true);
        if (cudaLaunchKernelEx(&L.cfg, k_match, V, det_tlwh, det_count, out_det_track_id) != cudaSuccess) return DD_ERR_CUDA;
    }
    if (ev) cudaEventRecord(ev[4], st);
    {
        DDLaunch L(items_to_blocks((long long)V.S * V.D), DD_WARPS * 32, 0, st, true);
        if (cudaLaunchKernelEx(&L.cfg, k_apply, V, det_conf, det_label) != cudaSuccess) return DD_ERR_CUDA;
    }
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
                               void* const* host_events6, void* gallery_wait, void* gallery_done) {
    if (!host_events6) return DD_ERR_INVALID;
    cudaEvent_t ev[6];
    for (int i = 0; i < 6; ++i) ev[i] = (cudaEvent_t)host_events6[i];
    return dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count,
                          out_det_track_id, (cudaStream_t)stream, ev, (cudaEvent_t)gallery_wait,
                          (cudaEvent_t)gallery_done);
}

int dd_tuning_set(int32_t key, int32_t value) {
    if (key == 0 && value >= 0 && value <= 3) { g_gate_impl = value; return DD_OK; }
    if (key == 1 && value >= 1 && value <= 16) { g_cosine_ctas_per_sm = value; return DD_OK; }
    if (key == 2 && (value == 0 || value == 1)) { g_small_priority = value; return DD_OK; }
    if (key == 3 && (value == 0 || value == 1)) { g_gallery_streaming = value; return DD_OK; }
    if (key == 4 && (value == 0 || value == 1)) { g_pdl = value; return DD_OK; }
    if (key == 5 && value >= -1 && value <= 2) { g_match_cta = value; return DD_OK; }
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

int dd_tracker_tick_chained(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                            const float* det_conf, const int32_t* det_label, const float* det_feat,
                            const int32_t* det_count, int32_t* out_det_track_id, const double* line,
                            int line_per_stream, int64_t* out_counts, void* gallery_wait, void* gallery_done,
                            void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!line) return DD_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    rc = dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count, out_det_track_id, st, nullptr,
                        (cudaEvent_t)gallery_wait, (cudaEvent_t)gallery_done, /*with_predict=*/true);
    if (rc != DD_OK) return rc;
    {
        DDLaunch L(warps_to_blocks(V.S), DD_WARPS * 32, 0, st, true);
        if (cudaLaunchKernelEx(&L.cfg, k_countline, V, line, line_per_stream) != cudaSuccess) return DD_ERR_CUDA;
    }
    if (out_counts) {
        DDLaunch L(V.C * 4, 256, 0, st, true);
        if (cudaLaunchKernelEx(&L.cfg, k_count_reduce, (const long long*)V.counts, V.S, V.C * 4, (long long*)out_counts) != cudaSuccess)
            return DD_ERR_CUDA;
    }
    return DD_OK;
}

int dd_tracker_tick(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                    const float* det_conf, const int32_t* det_label, const float* det_feat,
                    const int32_t* det_count, int32_t* out_det_track_id, const double* line,
                    int line_per_stream, int64_t* out_counts, void* stream) {
    return dd_tracker_tick_chained(state, cfg, det_tlwh, det_conf, det_label, det_feat, det_count, out_det_track_id,
                                   line, line_per_stream, out_counts, nullptr, nullptr, stream);
}

int dd_tracker_tick_ragged(void* state, const dd_tracker_config* cfg, const void* blob, int64_t off_tlwh,
                           int64_t off_conf, int64_t off_label, int64_t off_feat, double* det_tlwh, float* det_conf,
                           int32_t* det_label, int32_t* det_count, int32_t* out_det_track_id, const double* line,
                           int line_per_stream, int64_t* out_counts, void* stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    if (!line || !blob) return DD_ERR_INVALID;
    if ((off_tlwh & 7) || (off_conf & 3) || (off_label & 3) || (off_feat & 15) || ((uintptr_t)blob & 15)) return DD_ERR_INVALID;
    DDRagged R;
    R.blob = (const unsigned char*)blob;
    R.off_tlwh = off_tlwh; R.off_conf = off_conf; R.off_label = off_label; R.off_feat = off_feat;
    R.det_tlwh = det_tlwh; R.det_conf = det_conf; R.det_label = det_label; R.det_count = det_count;
    cudaStream_t st = (cudaStream_t)stream;
    rc = dd_update_impl(state, cfg, det_tlwh, det_conf, det_label, nullptr, det_count, out_det_track_id, st, nullptr,
                        nullptr, nullptr, /*with_predict=*/true, &R);
    if (rc != DD_OK) return rc;
    {
        DDLaunch L(warps_to_blocks(V.S), DD_WARPS * 32, 0, st, true);
        if (cudaLaunchKernelEx(&L.cfg, k_countline, V, line, line_per_stream) != cudaSuccess) return DD_ERR_CUDA;
    }
    if (out_counts) {
        DDLaunch L(V.C * 4, 256, 0, st, true);
        if (cudaLaunchKernelEx(&L.cfg, k_count_reduce, (const long long*)V.counts, V.S, V.C * 4, (long long*)out_counts) != cudaSuccess)
            return DD_ERR_CUDA;
    }
    return DD_OK;
}

int dd_unpack_detections(const void* blob, int32_t n_streams, int32_t max_dets, int64_t off_tlwh, int64_t off_conf,
                         int64_t off_label, int64_t off_feat, double* det_tlwh, float* det_conf,
                         int32_t* det_label, float* det_feat, int32_t* det_count, void* stream) {
    if (!blob || !det_tlwh || !det_conf || !det_label || !det_feat || !det_count) return DD_ERR_INVALID;
    if (n_streams <= 0 || max_dets <= 0) return DD_ERR_INVALID;
    if ((off_tlwh & 7) || (off_conf & 3) || (off_label & 3) || (off_feat & 15) || ((uintptr_t)blob & 15)) return DD_ERR_INVALID;
    k_unpack<<<warps_to_blocks((long long)n_streams * max_dets), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)blob, n_streams, max_dets, off_tlwh, off_conf, off_label, off_feat, det_tlwh, det_conf,
        det_label, det_feat, det_count);
    DD_CHECK_LAUNCH();
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
