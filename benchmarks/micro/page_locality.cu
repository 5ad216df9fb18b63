// Does the gallery stream's access pattern -- 4 KB half pages scattered over a multi-GB pool -- cap its bandwidth?
// Same data path as k_gallery_stream's producer / mma warps (cp.async.bulk into a 4-stage shared-memory ring per warp
// pair, mbarrier hand-over, consumer reads the stage with LDS.128 and releases it), on three page orders:
//   random pages | random runs of R consecutive pages | linear.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o page_locality page_locality.cu && ./page_locality
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#define PAGE 4096
__device__ __forceinline__ void mma_f16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// pairs of warps (producer, consumer) per CTA; each pair owns `stages` ring slots; pages[] lists the page ids to read,
// pair q of the grid takes entries q, q + n_pairs, ...
template <int MODE>
__global__ void k_stream3(const char* pool, const char* pool2, size_t pool2_rows, const int* pages, int n, int stages,
                          unsigned long long* sink) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = warp / 3, role = warp % 3;
    const int pairs_per_cta = blockDim.x / 96;
    char* ring = smem + (size_t)pair * (stages * PAGE + 1024 + 256);
    char* small = ring + stages * PAGE;
    unsigned long long* full = (unsigned long long*)(small + 1024);
    unsigned long long* empty = full + stages;
    unsigned long long* sfull = empty + stages;
    if (role == 0 && lane == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(sfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int q = blockIdx.x * pairs_per_cta + pair, nq = gridDim.x * pairs_per_cta;
    int st = 0;
    if (role == 0) {
        unsigned ph = ~0u;
        int k = 0;
        for (int i = q; i < n; i += nq, ++k) {
            if ((MODE & 4) && k % 7 == 0 && lane < 4)       // 4 small copies per 7 pages (nobody waits for them here)
                bulk(small + lane * 256, pool2 + ((size_t)(pages[i] * 2654435761u + lane) % pool2_rows) * 512, 256, sfull);
            mbar_wait(empty + st, (ph >> st) & 1u);
            ph ^= 1u << st;
            if (lane == 0) {
                mbar_expect(full + st, PAGE);
                bulk(ring + st * PAGE, pool + (size_t)pages[i] * PAGE, PAGE, full + st);
            }
            st = st + 1 == stages ? 0 : st + 1;
        }
    } else if (role == 1) {
        unsigned ph = 0u;
        float acc = 0.f;
        for (int i = q; i < n; i += nq) {
            mbar_wait(full + st, (ph >> st) & 1u);
            ph ^= 1u << st;
            const uint4* pg = (const uint4*)(ring + st * PAGE);
            uint4 ga[4], gb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { ga[j] = pg[j * 64 + lane]; gb[j] = pg[j * 64 + 32 + lane]; }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
            st = st + 1 == stages ? 0 : st + 1;
            if (MODE & 1) {
                float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_f16(ca, ga[j].x, gb[j].x, ga[j].y, gb[j].y, ga[j].x, ga[j].y);
                    mma_f16(cb, ga[j].z, gb[j].z, ga[j].w, gb[j].w, ga[j].z, ga[j].w);
                }
                acc += ca[0] + cb[0] + ca[1] + cb[1] + ca[2] + cb[2] + ca[3] + cb[3];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc += __uint_as_float((ga[j].x ^ gb[j].y) & 0x3fffffffu);
            }
        }
        if (acc == 1.2345f) sink[0] = 1;
    } else {
        if (MODE & 2) {           // "checker": per 7 pages two dependent rounds of 4 random 512-byte rows
            float acc = 0.f;
            unsigned h = q * 2654435761u + 12345u;
            for (int i = q; i < n; i += nq * 7) {
                for (int r = 0; r < 2; ++r) {
                    float4 a[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        h = h * 1664525u + 1013904223u + (unsigned)(acc != 7.f);
                        a[k] = ((const float4*)(pool2 + ((size_t)h % pool2_rows) * 512))[lane];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc += a[k].x + a[k].y + a[k].z + a[k].w;
                    acc = __shfl_xor_sync(0xffffffffu, acc, 1) + 1.f;      // next round depends on this one
                }
            }
            if (acc == 1.2345f) sink[1] = 1;
        }
    }
}

__global__ void k_stream(const char* pool, const int* pages, int n, int stages, unsigned long long* sink) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = warp >> 1, role = warp & 1;
    const int pairs_per_cta = blockDim.x >> 6;
    char* ring = smem + (size_t)pair * (stages * PAGE + 256);
    unsigned long long* full = (unsigned long long*)(ring + stages * PAGE);
    unsigned long long* empty = full + stages;
    if (role == 0 && lane == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int q = blockIdx.x * pairs_per_cta + pair, nq = gridDim.x * pairs_per_cta;
    int st = 0;
    if (role == 0) {
        unsigned ph = ~0u;
        for (int i = q; i < n; i += nq) {
            mbar_wait(empty + st, (ph >> st) & 1u);
            ph ^= 1u << st;
            if (lane == 0) {
                mbar_expect(full + st, PAGE);
                bulk(ring + st * PAGE, pool + (size_t)pages[i] * PAGE, PAGE, full + st);
            }
            st = st + 1 == stages ? 0 : st + 1;
        }
    } else {
        unsigned ph = 0u;
        unsigned acc = 0;
        for (int i = q; i < n; i += nq) {
            mbar_wait(full + st, (ph >> st) & 1u);
            ph ^= 1u << st;
            const uint4* pg = (const uint4*)(ring + st * PAGE);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const uint4 v = pg[j * 32 + lane]; acc += v.x ^ v.y ^ v.z ^ v.w; }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
            st = st + 1 == stages ? 0 : st + 1;
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

int main(int argc, char** argv) {
    const size_t pool_pages = (size_t)(argc > 1 ? atof(argv[1]) : 3.0) * (1ull << 30) / PAGE;     // GB of pool
    const int n = 330000;                      // pages per launch (1.35 GB, like one C3 tick)
    char* pool;
    cudaMalloc(&pool, pool_pages * PAGE);
    cudaMemset(pool, 1, pool_pages * PAGE);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::mt19937_64 rng(1);
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    int* d_pages;
    cudaMalloc(&d_pages, n * sizeof(int));
    const int runs[] = {1, 2, 4, 7, 16, 64, 0};            // 0 = linear
    for (int pairs : {7, 10}) {
        for (int stages : {4, 6}) {
            const size_t smem = (size_t)pairs * (stages * PAGE + 256);
            if (smem > 227 * 1024) continue;
            cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int R : runs) {
                std::vector<int> pages(n);
                if (R == 0) {
                    for (int i = 0; i < n; ++i) pages[i] = i;
                } else {
                    // random runs of R consecutive pages; a pair reads entries q, q + nq, ...: lay the list out so that one
                    // pair walks a run front to back (what a producer warp does with one track's gallery)
                    const int nq = sms * pairs;
                    std::vector<int> starts((n + R - 1) / R);
                    for (auto& s : starts) s = (int)(rng() % (pool_pages / R)) * R;
                    for (int i = 0; i < n; ++i) {
                        const int q = i % nq, k = i / nq;              // k-th page of pair q
                        const long long idx = (long long)q * ((n + nq - 1) / nq) + k;      // position in pair q's own sequence
                        pages[i] = starts[(idx / R) % starts.size()] + (int)(idx % R);
                    }
                }
                cudaMemcpy(d_pages, pages.data(), n * sizeof(int), cudaMemcpyHostToDevice);
                cudaEvent_t a, b;
                cudaEventCreate(&a); cudaEventCreate(&b);
                for (int w = 0; w < 3; ++w) k_stream<<<sms, pairs * 64, smem>>>(pool, d_pages, n, stages, sink);
                cudaEventRecord(a);
                const int it = 10;
                for (int w = 0; w < it; ++w) k_stream<<<sms, pairs * 64, smem>>>(pool, d_pages, n, stages, sink);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                ms /= it;
                printf("{\"pairs_per_sm\": %d, \"stages\": %d, \"run_pages\": %d, \"ms\": %.4f, \"GBps\": %.1f, \"err\": \"%s\"}\n", pairs, stages, R,
                       ms, (double)n * PAGE / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    // ---- the same ring with the gallery stream's other ingredients added one by one (random pages, 7 triples x 4 stages)
    {
        const size_t pool2_rows = (size_t)2 * (1ull << 30) / 512;
        char* pool2;
        cudaMalloc(&pool2, pool2_rows * 512);
        cudaMemset(pool2, 1, pool2_rows * 512);
        std::vector<int> pages(n);
        for (auto& p : pages) p = (int)(rng() % pool_pages);
        cudaMemcpy(d_pages, pages.data(), n * sizeof(int), cudaMemcpyHostToDevice);
        const int triples = 7, stages = 4;
        const size_t smem = (size_t)triples * (stages * PAGE + 1024 + 256);
        auto run = [&](auto kern, int mode) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            for (int w = 0; w < 3; ++w) kern<<<sms, triples * 96, smem>>>(pool, pool2, pool2_rows, d_pages, n, stages, sink);
            cudaEventRecord(a);
            const int it = 10;
            for (int w = 0; w < it; ++w) kern<<<sms, triples * 96, smem>>>(pool, pool2, pool2_rows, d_pages, n, stages, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            ms /= it;
            printf("{\"triples\": 7, \"stages\": 4, \"mode\": %d, \"ms\": %.4f, \"page_GBps\": %.1f, \"err\": \"%s\", \"modes\": \"1 = 8 HMMA per page, 2 = third warp: 2 dependent rounds of 4 random 512 B rows per 7 pages, 4 = 4 small bulk copies per 7 pages\"}\n",
                   mode, ms, (double)n * PAGE / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        };
        run(k_stream3<0>, 0); run(k_stream3<1>, 1); run(k_stream3<2>, 2); run(k_stream3<3>, 3); run(k_stream3<4>, 4); run(k_stream3<7>, 7);
    }
    return 0;
}
