// dd_gallery.cuh -- the gallery kernels: min cosine distance between the gate-passing detections of a track and the
// track's feature gallery (nn_matching.py:31-54,78-96 under tracker.py:97-105), over paged galleries.
//
//   k_cosine          exact f32 pass, one warp per track index (gallery_impl = 1; defines the result)
//   k_gallery_stream  half-precision pre-pass on tensor cores + exact re-check, as a producer / mma / checker warp
//                     pipeline over TMA-fed shared-memory rings (gallery_impl = 0, the default)
//
// The half pre-pass ("half the bytes, the same bits").  The exact pass is bound by the 512 B per gallery row it must
// read, so every gallery page has a round-to-nearest HALF shadow (256 B per row) and the kernels stream that:
//   a(r, n) = tensor-core dot (mma.sync m16n8k16, f16 inputs, f32 accumulate) of half row r and half query n,
//   e(r, n) = the f32 value the exact pass computes (4 FMAs per lane + the 16-8-4-2-1 butterfly).
// |a - e| <= E := 1.1e-3 for unit vectors (2^-10 from rounding both operands to half, Cauchy-Schwarz; 1e-4 of
// slack for the tensor-core accumulation and 1e-5 for the f32 pass itself).  With m = max_r a(r, n) the row
// r* that maximises e satisfies a(r*, n) >= m - 2E, so the exact maximum is the maximum of e over the rows with
// a >= m - 2E -- typically one or two rows, read from the f32 page and evaluated with exactly the arithmetic of
// the exact pass.  The cost matrix is therefore bit-identical, whatever the data; only the number of re-checked
// rows (speed) depends on it.  Galleries longer than DD_GS_BLOCK_ROWS are processed block by block with the running
// maximum m' <= m of the blocks so far: the window only gets wider, never wrong.
// Fragment trick: a dot product does not care about the order of its terms, so lane 4 g + t feeds the mma with the
// eight consecutive halves of one 16-byte chunk (chunk t + 4 j of rows g and g + 8, j = 0..3) and takes the B operand
// from the same chunk of query g; half pages are stored in that order (dd_half_chunk_off), so every load instruction
// of a warp covers 512 contiguous bytes.  Up to 8 gate-passing detections share one pass over the gallery.
#pragma once
#include "dd_tracker_bodies.cuh"

#define DD_WARPS 4
#define DD_H_WINDOW 2.2e-3f

__global__ void __launch_bounds__(DD_WARPS * 32, 7)
k_cosine(const DDView V, const DDTickArgs A) {
    const int* __restrict__ det_count = DD_ARG(det_count);
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= V.S * V.T) return;
    WarpG g;
    DDDirectPass<WarpG> pass;
    dd_cosine_track(g, V, w / V.T, w % V.T, det_count, pass);
}

__device__ __forceinline__ void dd_mma_f16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                           unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- k_gallery_stream: the half pre-pass as a three-stage warp pipeline (gallery_impl = 0, the default) -----------
// A per-warp kernel (round 1's k_cosine_h) keeps its in-flight gallery bytes in registers and serialises, per track, a
// chain of dependent latencies (claim -> descriptor -> gate word -> queries -> rows ... -> candidate list -> f32 re-check
// rows).  Here every link of that chain is its own warp and the bytes in flight live in shared memory.  A CTA is a set of
// warp TRIPLES, one CTA per SM:
//   producer P   walks the work list (self-contained records written by k_gate, claimed two entries ahead) and, for
//                every job = (track, group of <= 8 gate-passing detections, block of <= 128 gallery rows), posts a job
//                header + the detections' half rows, then feeds the triple's ring of 4 KB stages with ONE bulk
//                asynchronous copy per gallery page (cp.async.bulk, completion on the stage's mbarrier; a ragged last
//                page is copied row-block by row-block so that no byte past the gallery's end is read).  It runs across
//                job and track boundaries, as far ahead as the ring allows.
//   mma warp M   waits for a stage, takes its mma fragments with conflict-free 16-byte shared loads (pages are stored
//                in fragment order), releases the stage at once, does the 8 mma of the page and keeps the approximate
//                dots of the whole block in REGISTERS; after the block it names the window candidates by ballot and
//                posts them to the checker.  It never touches global memory.
//   checker C    evaluates the candidates from the f32 pages with the exact pass's arithmetic (8 rows in flight),
//                keeps the exact maximum per detection and writes the cost entries.  Its dependent global reads stall
//                nobody but itself.
#include "dd_tma.cuh"

#define DD_GS_BLOCK_ROWS 128                 // gallery rows per job (8 pages): the mma warp keeps 8 x 4 dots per lane
#define DD_GS_HDR_INTS 32                    // job header: 0 slotg 1 stream 2 row0 3 nrows 4 nq 5 flags | 8.. cj[8] | 16.. pid[8]
#ifndef DD_GS_MQ
#define DD_GS_MQ 2                           // checker messages in flight per triple: the checker's service time has a long
#endif                                       // tail (dependent global reads), a shallow queue would stall the mma warp behind it
#define DD_GS_MSG_CAP (DD_GS_MQ > 2 ? 112 : 240)   // candidates a message can list; a block with more (near-identical rows:
                                             // legal, never seen outside the adversarial tests) is re-evaluated in full
#define DD_GS_MSG_BYTES (DD_GS_HDR_INTS * 4 + DD_GS_MSG_CAP * 2)    // checker message: header (32 ints; [6] = candidates) + list (u16)
// SKIP (template parameter of the role bodies): timing experiments only -- the product kernel is SKIP = 0; variants are
// compiled under -DDD_GS_VARIANTS and replayed on a finished tick's work list by dd_gallery_replay (wrong costs, nothing
// reads them): 1 = checker evaluates nothing, 2 = no candidate listing, 4 = no fragment loads / mma, 8 = one query row per
// job, 16 = a ragged last page is copied whole.
#define DD_GS_FIRST 1                        // first job / message of a detection group: reset the running maxima
#define DD_GS_LAST 2                         // last one: the cost entries are final
#define DD_GS_STOP 4

__host__ __device__ inline size_t dd_gs_triple_bytes(int stages) {
    const size_t b = (size_t)stages * DD_PAGE_F16_BYTES            // ring
                     + 2 * 8 * 256                                 // query half rows, double-buffered
                     + 2 * DD_GS_HDR_INTS * 4                      // job headers, double-buffered
                     + DD_GS_MQ * DD_GS_MSG_BYTES                  // checker messages
                     + DD_GS_BLOCK_ROWS * 8 * 2                    // approximate dots of the block being streamed (half)
                     + (size_t)(2 * stages + 4 + 2 * DD_GS_MQ) * 8; // mbarriers: full / empty [stages], hfull hfree [2], mfull mfree [MQ]
    return (b + 127) & ~(size_t)127;
}

__device__ __forceinline__ void dd_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dd_smem_u32(bar)) : "memory");
}

#ifdef DD_GS_VARIANTS
// role-level cycle accounting of the variant builds (benchmarks/gallery_variants.py): cycles every role spent inside
// its mbarrier waits, summed over all triples.  0 producer total, 1 wait empty, 2 wait hfree | 4 mma total, 5 wait full,
// 6 wait hfull, 7 wait mfree | 8 checker total, 9 wait mfull
__device__ unsigned long long dd_gs_prof[16];
#define DD_GS_PROF_DECL unsigned long long pw_[4] = {0, 0, 0, 0}; const long long pt0_ = clock64();
#define DD_GS_WAIT(k, bar, par) do { const long long c0_ = clock64(); dd_mbar_wait(bar, par); pw_[k] += (unsigned long long)(clock64() - c0_); } while (0)
#define DD_GS_PROF_FLUSH(base) do { if ((threadIdx.x & 31) == 0) { atomicAdd(dd_gs_prof + (base), (unsigned long long)(clock64() - pt0_)); \
        for (int k_ = 1; k_ < 4; ++k_) atomicAdd(dd_gs_prof + (base) + k_, pw_[k_]); } } while (0)
#else
#define DD_GS_PROF_DECL
#define DD_GS_WAIT(k, bar, par) dd_mbar_wait(bar, par)
#define DD_GS_PROF_FLUSH(base)
#endif

struct DDTripleSmem {
    char* ring;
    char* qbuf;
    int* hdr;
    char* msg;
    unsigned* approx;            // [DD_GS_BLOCK_ROWS][4] half2: dots of row r with detections 2 t, 2 t + 1
    unsigned long long *full, *empty, *hfull, *hfree, *mfull, *mfree;
};
__device__ __forceinline__ void dd_gs_carve(char* base, int stages, DDTripleSmem& P) {
    P.ring = base;
    P.qbuf = P.ring + (size_t)stages * DD_PAGE_F16_BYTES;
    P.hdr = (int*)(P.qbuf + 2 * 8 * 256);
    P.msg = (char*)(P.hdr + 2 * DD_GS_HDR_INTS);
    P.approx = (unsigned*)(P.msg + DD_GS_MQ * DD_GS_MSG_BYTES);
    P.full = (unsigned long long*)((char*)P.approx + DD_GS_BLOCK_ROWS * 8 * 2);
    P.empty = P.full + stages;
    P.hfull = P.empty + stages;
    P.hfree = P.hfull + 2;
    P.mfull = P.hfree + 2;
    P.mfree = P.mfull + DD_GS_MQ;
}

// quota: work-list entries this producer may claim (even; INT_MAX = until the list is empty).  With the one-triple-
// per-CTA launch every CTA streams a bounded share and exits, so that its shared memory becomes available to
// whatever kernel is waiting (the quotas of all CTAs together cover the list).
template <int SKIP>
__device__ __forceinline__ void dd_gs_producer(const DDView& V, const DDTripleSmem& P, int stages, int quota) {
    const int lane = threadIdx.x & 31;
    const int n = V.work_ctl[0];
    int claimed = 0;
    DD_GS_PROF_DECL
    int st = 0;                  // next ring stage
    unsigned ephase = ~0u;       // bit s: parity to wait for on empty[s] (a fresh barrier passes a wait on parity 1)
    int hb = 0;
    unsigned hphase = 3u;        // same for hfree[0..1]
    // claims are made two work-list entries at a time (one atomic per pair of tracks) and one pair ahead: the
    // claim made when pair k starts is broadcast when it ends
    int raw = n;
    if (lane == 0) raw = atomicAdd(V.work_ctl + 32, 2);
    claimed += 2;
    int i_cur = __shfl_sync(0xffffffffu, raw, 0);          // i_cur, i_cur + 1: the pair being streamed
    raw = n;
    if (claimed < quota) {
        if (lane == 0) raw = atomicAdd(V.work_ctl + 32, 2);
        claimed += 2;
    }
    int i_nxt = __shfl_sync(0xffffffffu, raw, 0);          // first entry of the next pair
    int odd = 0;                                           // 0: streaming the pair's first entry, 1: its second
    int w_cur = i_cur < n ? V.work_rec[(size_t)i_cur * 16 + (lane & 15)] : 0;
    while (i_cur < n) {
        const int i_after = odd ? i_nxt : i_cur + 1;                                         // the entry streamed next
        const int w_nxt = i_after < n ? V.work_rec[(size_t)i_after * 16 + (lane & 15)] : 0;  // used one entry later
        if (!odd) {                                                                          // broadcast two entries later
            raw = n;
            if (claimed < quota) {
                if (lane == 0) raw = atomicAdd(V.work_ctl + 32, 2);
                claimed += 2;
            }
        }
        const int slotg = __shfl_sync(0xffffffffu, w_cur, 0), s = __shfl_sync(0xffffffffu, w_cur, 1);
        const int glen = __shfl_sync(0xffffffffu, w_cur, 2), np = __shfl_sync(0xffffffffu, w_cur, 4);
        const unsigned gw0 = (unsigned)__shfl_sync(0xffffffffu, w_cur, 5), gw1 = (unsigned)__shfl_sync(0xffffffffu, w_cur, 6);
        const int nd = __shfl_sync(0xffffffffu, w_cur, 7);
        const int nblocks = glen > 0 ? (glen + DD_GS_BLOCK_ROWS - 1) / DD_GS_BLOCK_ROWS : 1;
        int base = 0, wi = 0;
        unsigned word = gw0;
        for (;;) {
            // ---- next group of <= 8 gate-passing detections; lane q keeps detection q
            int nq = 0, cjl = 0;
            while (nq < 8) {
                if (!word) {
                    base += 32;
                    ++wi;
                    if (base >= nd) break;
                    word = wi == 1 ? gw1 : V.gate[(size_t)slotg * V.DW + wi];
                    continue;
                }
                if (lane == nq) cjl = base + dd_ctz(word);
                ++nq;
                word &= word - 1;
            }
            if (nq == 0) break;
            for (int b = 0; b < nblocks; ++b) {
                const int row0 = b * DD_GS_BLOCK_ROWS;
                const int nrows = glen > 0 ? dd_imin(DD_GS_BLOCK_ROWS, glen - row0) : 0;
                const int npg = (nrows + 15) >> 4;
                // page ids of this block: lane k keeps page b * 8 + k (the record carries the first 8)
                int pidl = 0;
                {
                    const int k = b * 8 + (lane & 7);
                    const int from_rec = __shfl_sync(0xffffffffu, w_cur, 8 + (lane & 7));
                    pidl = b == 0 ? from_rec : (k < np ? V.ptab[(size_t)slotg * V.PT + k] : 0);
                }
                // ---- job header + query rows
                DD_GS_WAIT(2, P.hfree + hb, (hphase >> hb) & 1u);
                hphase ^= 1u << hb;
                int* H = P.hdr + hb * DD_GS_HDR_INTS;
                if (lane == 0) {
                    H[0] = slotg; H[1] = s; H[2] = row0; H[3] = nrows; H[4] = nq;
                    H[5] = (b == 0 ? DD_GS_FIRST : 0) | (b == nblocks - 1 ? DD_GS_LAST : 0);
                }
                if (lane < 8) { H[8 + lane] = cjl; H[16 + lane] = pidl; }
                __syncwarp();
                if (lane == 0) dd_mbar_expect_tx(P.hfull + hb, (SKIP & 8) ? 256u : (unsigned)nq * 256u);
                __syncwarp();
                if (lane < ((SKIP & 8) ? 1 : nq))
                    dd_bulk_g2s(P.qbuf + hb * 2048 + lane * 256, V.det_feath + ((size_t)s * V.D + cjl) * DD_FEAT_DIM, 256u,
                                P.hfull + hb);
                hb ^= 1;
                // ---- the block's pages
                for (int p = 0; p < npg; ++p) {
                    DD_GS_WAIT(1, P.empty + st, (ephase >> st) & 1u);
                    ephase ^= 1u << st;
                    const int valid = dd_imin(16, nrows - p * 16);
                    const char* src = dd_page_f16(V, __shfl_sync(0xffffffffu, pidl, p));
                    char* dst = P.ring + (size_t)st * DD_PAGE_F16_BYTES;
                    if (valid == 16 || (SKIP & 16)) {
                        if (lane == 0) {
                            dd_mbar_expect_tx(P.full + st, DD_PAGE_F16_BYTES);
                            dd_bulk_g2s(dst, src, DD_PAGE_F16_BYTES, P.full + st);
                        }
                    } else {
                        // ragged last page: 512-byte block l = (j, h) holds chunk group j of rows 8 h .. 8 h + 7, 64 B per row
                        if (lane == 0) dd_mbar_expect_tx(P.full + st, (unsigned)valid * 256u);
                        __syncwarp();
                        if (lane < 8) {
                            const int rows = dd_imin(8, dd_imax(0, valid - 8 * (lane & 1)));
                            if (rows > 0) dd_bulk_g2s(dst + lane * 512, src + lane * 512, (unsigned)rows * 64u, P.full + st);
                        }
                    }
                    st = st + 1 == stages ? 0 : st + 1;
                }
            }
        }
        i_cur = i_after;
        w_cur = w_nxt;
        if (odd) i_nxt = __shfl_sync(0xffffffffu, raw, 0);
        odd ^= 1;
    }
    // ---- stop job
    dd_mbar_wait(P.hfree + hb, (hphase >> hb) & 1u);
    if (lane == 0) {
        P.hdr[hb * DD_GS_HDR_INTS + 5] = DD_GS_STOP;
        dd_mbar_arrive(P.hfull + hb);
    }
    DD_GS_PROF_FLUSH(0);
}

template <int SKIP>
__device__ __forceinline__ void dd_gs_mma(const DDTripleSmem& P, int stages) {
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2, tq = lane & 3;
    int st = 0, hb = 0, mb = 0;
    unsigned fphase = 0u;        // bit s: parity to wait for on full[s]
    unsigned hphase = 0u;        // hfull[0..1]
    unsigned mphase = ~0u;       // mfree[0..MQ-1] (fresh barriers pass)
    float run0 = -3.0e38f, run1 = -3.0e38f;     // approximate maxima so far of detections 2 tq, 2 tq + 1 (over the group's blocks)
    DD_GS_PROF_DECL
    for (;;) {
        DD_GS_WAIT(2, P.hfull + hb, (hphase >> hb) & 1u);
        hphase ^= 1u << hb;
        const int hw = P.hdr[hb * DD_GS_HDR_INTS + lane];      // lane l keeps header word l
        const int flags = __shfl_sync(0xffffffffu, hw, 5);
        if (flags & DD_GS_STOP) {
            dd_mbar_wait(P.mfree + mb, (mphase >> mb) & 1u);
            if (lane == 0) {
                ((int*)(P.msg + mb * DD_GS_MSG_BYTES))[5] = DD_GS_STOP;
                dd_mbar_arrive(P.mfull + mb);
            }
            DD_GS_PROF_FLUSH(4);
            break;
        }
        const int nrows = __shfl_sync(0xffffffffu, hw, 3), nq = __shfl_sync(0xffffffffu, hw, 4);
        uint4 qb[4];
        {
            const uint4* qh = (const uint4*)(P.qbuf + hb * 2048 + gq * 256);
#pragma unroll
            for (int j = 0; j < 4; ++j) qb[j] = gq < nq ? qh[tq + 4 * j] : make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        if (lane == 0) dd_mbar_arrive(P.hfree + hb);           // header and queries are in registers
        hb ^= 1;
        if (flags & DD_GS_FIRST) { run0 = -3.0e38f; run1 = -3.0e38f; }
        const int npg = (nrows + 15) >> 4;
        float mx0 = -3.0e38f, mx1 = -3.0e38f;
        for (int p = 0; p < npg; ++p) {                        // deliberately not unrolled: the loop must stay in the L0 i-cache
            DD_GS_WAIT(1, P.full + st, (fphase >> st) & 1u);
            fphase ^= 1u << st;
            const uint4* pg = (const uint4*)(P.ring + (size_t)st * DD_PAGE_F16_BYTES);
            uint4 ga[4], gb[4];
            if (SKIP & 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { ga[j] = make_uint4(lane, p, j, 1); gb[j] = ga[j]; }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { ga[j] = pg[j * 64 + lane]; gb[j] = pg[j * 64 + 32 + lane]; }
            }
            __syncwarp();
            if (lane == 0) dd_mbar_arrive(P.empty + st);       // the stage is free again: its bytes are in registers
            st = st + 1 == stages ? 0 : st + 1;
            // two independent accumulator chains (any summation order satisfies the window bound)
            float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
            if (SKIP & 4) {
                ca[0] = __uint_as_float(ga[0].x & 0xffu) * 1e-9f;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dd_mma_f16(ca, ga[j].x, gb[j].x, ga[j].y, gb[j].y, qb[j].x, qb[j].y);
                    dd_mma_f16(cb, ga[j].z, gb[j].z, ga[j].w, gb[j].w, qb[j].z, qb[j].w);
                }
            }
            float c0 = ca[0] + cb[0], c1 = ca[1] + cb[1], c2 = ca[2] + cb[2], c3 = ca[3] + cb[3];
            const int ra = p * 16 + gq, rb = ra + 8;
            if (ra >= nrows) { c0 = -3.0e38f; c1 = -3.0e38f; }     // rows past the end (stale stage bytes) never win
            if (rb >= nrows) { c2 = -3.0e38f; c3 = -3.0e38f; }
            // parked as half2 (every kilobyte of shared memory here is a matching warp of another chunk that fits on the
            // SM): the candidate test below widens the window by the rounding error, so the list stays a superset
            {
                const __half2 ha = __floats2half2_rn(c0, c1), hb = __floats2half2_rn(c2, c3);
                P.approx[ra * 4 + tq] = *(const unsigned*)&ha;
                P.approx[rb * 4 + tq] = *(const unsigned*)&hb;
            }
            mx0 = fmaxf(mx0, fmaxf(c0, c2));
            mx1 = fmaxf(mx1, fmaxf(c1, c3));
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        }
        run0 = fmaxf(run0, mx0);
        run1 = fmaxf(run1, mx1);
        // a detection column this lane holds no real query for can never list a candidate
        // the parked dots are rounded to half: |half(a) - a| <= 2^-11 |a| < 5e-4 for |a| <= 1.001
        const float thr0 = 2 * tq < nq ? run0 - DD_H_WINDOW - 5.0e-4f : 3.0e38f;
        const float thr1 = 2 * tq + 1 < nq ? run1 - DD_H_WINDOW - 5.0e-4f : 3.0e38f;
        // ---- candidates -> checker message: every lane re-reads the dots it wrote and appends its own candidates
        // (a shared-memory counter hands out list positions; the order of the list does not matter to a maximum)
        DD_GS_WAIT(3, P.mfree + mb, (mphase >> mb) & 1u);
        mphase ^= 1u << mb;
        int* M = (int*)(P.msg + mb * DD_GS_MSG_BYTES);
        unsigned short* mc = (unsigned short*)(M + DD_GS_HDR_INTS);
        M[lane] = lane == 6 ? 0 : hw;                          // word 6 = candidate counter
        __syncwarp();
        for (int p = 0; p < ((SKIP & 2) ? 0 : npg); ++p) {
            const int ra = p * 16 + gq, rb = ra + 8;
            const unsigned ua = P.approx[ra * 4 + tq], ub = P.approx[rb * 4 + tq];
            const float2 va = __half22float2(*(const __half2*)&ua);
            const float2 vb = __half22float2(*(const __half2*)&ub);
            // the counter keeps counting past the capacity: that is how the checker learns the list is incomplete
            const bool a0 = va.x >= thr0, a1 = va.y >= thr1, b0 = vb.x >= thr0, b1 = vb.y >= thr1;
            if (a0) { const int k = atomicAdd(M + 6, 1); if (k < DD_GS_MSG_CAP) mc[k] = (unsigned short)((ra << 3) | (2 * tq)); }
            if (a1) { const int k = atomicAdd(M + 6, 1); if (k < DD_GS_MSG_CAP) mc[k] = (unsigned short)((ra << 3) | (2 * tq + 1)); }
            if (b0) { const int k = atomicAdd(M + 6, 1); if (k < DD_GS_MSG_CAP) mc[k] = (unsigned short)((rb << 3) | (2 * tq)); }
            if (b1) { const int k = atomicAdd(M + 6, 1); if (k < DD_GS_MSG_CAP) mc[k] = (unsigned short)((rb << 3) | (2 * tq + 1)); }
        }
        __syncwarp();
        if (lane == 0) dd_mbar_arrive(P.mfull + mb);
        mb = mb + 1 == DD_GS_MQ ? 0 : mb + 1;
    }
}

template <int CW, int SKIP>
__device__ __forceinline__ void dd_gs_checker(const DDView& V, const DDTripleSmem& P) {
    const int lane = threadIdx.x & 31;
    int mb = 0;
    unsigned mphase = 0u;
    float best = -3.0e38f;       // lane n: exact maximum of detection n of the current group
    DD_GS_PROF_DECL
    for (;;) {
        DD_GS_WAIT(1, P.mfull + mb, (mphase >> mb) & 1u);
        mphase ^= 1u << mb;
        const int* M = (const int*)(P.msg + mb * DD_GS_MSG_BYTES);
        const unsigned short* mc = (const unsigned short*)(M + DD_GS_HDR_INTS);
        const int hw = M[lane];
        const int flags = __shfl_sync(0xffffffffu, hw, 5);
        if (flags & DD_GS_STOP) { DD_GS_PROF_FLUSH(8); break; }
        const int slotg = __shfl_sync(0xffffffffu, hw, 0), s = __shfl_sync(0xffffffffu, hw, 1);
        const int nq = __shfl_sync(0xffffffffu, hw, 4), listed = __shfl_sync(0xffffffffu, hw, 6);
        // more candidates than a message holds: every (row, detection) of the block is evaluated exactly instead
        const bool all = listed > DD_GS_MSG_CAP;
        const int ncand = (SKIP & 1) ? 0 : (all ? __shfl_sync(0xffffffffu, hw, 3) * nq : listed);
        if (flags & DD_GS_FIRST) best = -3.0e38f;
        const float4* qbase = (const float4*)(V.det_featn + (size_t)s * V.D * DD_FEAT_DIM);
        for (int c0 = 0; c0 < ncand; c0 += CW) {
            // exact values of CW listed (row, detection) entries: the exact pass's arithmetic, bit for bit
            float v[CW];
            float4 a[CW], q[CW];
            int qn[CW];
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                const int idx = dd_imin(c0 + k, ncand - 1);
                int e;
                if (all) e = ((idx / nq) << 3) | (idx % nq);
                else e = mc[idx];
                qn[k] = e & 7;
                const int row = e >> 3;
                const int pid = __shfl_sync(0xffffffffu, hw, 16 + (row >> 4));
                const int d = __shfl_sync(0xffffffffu, hw, 8 + qn[k]);
                a[k] = dd_page_f32(V, pid)[(size_t)(row & 15) * (DD_FEAT_DIM / 4) + lane];
                q[k] = qbase[(size_t)d * (DD_FEAT_DIM / 4) + lane];
            }
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                float p = dd_fmaf(a[k].x, q[k].x, 0.f);
                p = dd_fmaf(a[k].y, q[k].y, p);
                p = dd_fmaf(a[k].z, q[k].z, p);
                p = dd_fmaf(a[k].w, q[k].w, p);
                v[k] = p;
            }
            int nn = CW, o = 16;                        // transposing butterfly: the 16-8-4-2-1 summation tree of the exact pass
#pragma unroll
            for (; nn > 1; nn >>= 1, o >>= 1) {
                const bool up = (lane & o) != 0;
                const int half = nn >> 1;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const float send = up ? v[i] : v[i + half];
                    const float keep = up ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
#pragma unroll
            for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
            // lane L now holds the total of entry c0 + L / (32 / CW); lane n keeps the maximum of detection n
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                const float tot = __shfl_sync(0xffffffffu, v[0], (32 / CW) * k);
                if (c0 + k < ncand && qn[k] == lane) best = fmaxf(best, tot);
            }
        }
        const int dmine = __shfl_sync(0xffffffffu, hw, 8 + (lane & 7));
        if ((flags & DD_GS_LAST) && lane < nq) V.cost[(size_t)slotg * V.D + dmine] = dd_subf(1.0f, best);
        __syncwarp();
        if (lane == 0) dd_mbar_arrive(P.mfree + mb);
        mb = mb + 1 == DD_GS_MQ ? 0 : mb + 1;
    }
}

template <int SKIP>
__device__ __forceinline__ void dd_gs_body(const DDView& V, int stages, int bounded) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5;
    const int triple = warp / 3, role = warp - triple * 3;
    DDTripleSmem P;
    DD_TL_BEGIN(2);
    dd_gs_carve(smem + (size_t)triple * dd_gs_triple_bytes(stages), stages, P);
    if (role == 0 && (threadIdx.x & 31) == 0) {
        for (int i = 0; i < stages; ++i) { dd_mbar_init(P.full + i, 1); dd_mbar_init(P.empty + i, 1); }
        for (int i = 0; i < 2; ++i) { dd_mbar_init(P.hfull + i, 1); dd_mbar_init(P.hfree + i, 1); }
        for (int i = 0; i < DD_GS_MQ; ++i) { dd_mbar_init(P.mfull + i, 1); dd_mbar_init(P.mfree + i, 1); }
        dd_mbar_fence_init();
    }
    __syncthreads();
    int quota = 0x7ffffffe;
    if (bounded) {                       // even share of the list, rounded up: gridDim.x * quota >= entries
        const int n = V.work_ctl[0];
        quota = ((n + (int)gridDim.x - 1) / (int)gridDim.x + 1) & ~1;
        if (quota < 2) quota = 2;
    }
    if (role == 0) dd_gs_producer<SKIP>(V, P, stages, quota);
    else if (role == 1) dd_gs_mma<SKIP>(P, stages);
    else {
        dd_gs_checker<4, SKIP>(V, P);
        if (V.tl && (threadIdx.x & 31) == 0) atomicMax(DD_TL_SLOT(2) + 1, dd_globaltimer());   // the checker drains last
    }
}

// 72 registers (8 bytes of spills): the register file is four 16 K partitions, the CTA's 21 warps land 6 + 5 + 5 + 5,
// and at 72 every partition still has room for one 80-register matching warp of another chunk (at 80 the six-warp
// partition has not).
__global__ void __maxnreg__(72)
k_gallery_stream(const DDView V, int stages, int bounded) {
    dd_gs_body<0>(V, stages, bounded);
}

#ifdef DD_GS_VARIANTS
template <int SKIP>
__global__ void __maxnreg__(72)
k_gallery_stream_dbg(const DDView V, int stages, int bounded) {
    dd_gs_body<SKIP>(V, stages, bounded);
}
__global__ void k_reset_cursor(const DDView V) { V.work_ctl[32] = 0; }
#endif
