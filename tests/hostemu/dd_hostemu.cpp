// dd_hostemu.cpp -- TEST SCAFFOLDING ONLY (built by tests/hostemu/build.py, loaded only by tests/).
//
// Compiles the group-generic kernel bodies of deepdish_b200/csrc with HostG (one lane, trivial
// collectives) so that the serial logic of the CUDA path -- scipy-exact LSAP, CPython set order,
// cascade / lifecycle bookkeeping, count-line -- can be exercised against the oracle on a machine
// without a GPU.  It shares no code with oracle/, is never linked into libdeepdish_b200.so and is
// never imported by the deepdish_b200 package: the product path has no CPU fallback.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../deepdish_b200/csrc/dd_tracker_bodies.cuh"
#include "../../deepdish_b200/csrc/dd_detect_bodies.cuh"

extern "C" {

int ddh_tracker_layout_query(const dd_tracker_config* cfg, dd_tracker_layout* out) {
    return dd_layout_compute(cfg, out);
}

int ddh_tracker_init(void* state, const dd_tracker_config* cfg) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    dd_tracker_layout L;
    dd_layout_compute(cfg, &L);
    memset(state, 0, L.total_bytes);
    for (int s = 0; s < V.S; ++s) V.next_id[s] = 1;
    return DD_OK;
}

int ddh_tracker_predict(void* state, const dd_tracker_config* cfg) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    HostG g;
    for (int s = 0; s < V.S; ++s)
        for (int t = 0; t < V.T; ++t) dd_predict_track(g, V, s, t);
    return DD_OK;
}

int ddh_tracker_update(void* state, const dd_tracker_config* cfg, const double* det_tlwh,
                       const float* det_conf, const int32_t* det_label, const float* det_feat,
                       const int32_t* det_count, int32_t* out_det_track_id) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    HostG g;
    for (int s = 0; s < V.S; ++s)
        for (int d = 0; d < V.D; ++d) dd_prep_det(g, V, s, d, det_tlwh, det_feat, det_count);
    for (int s = 0; s < V.S; ++s)
        for (int t = 0; t < V.T; ++t) dd_gate_track(g, V, s, t, det_count);
    for (int s = 0; s < V.S; ++s)
        for (int t = 0; t < V.T; ++t) { DDDirectPass<HostG> pass; dd_cosine_track(g, V, s, t, det_count, pass); }
    std::vector<char> smem(dd_match_smem_bytes(V.T, V.D) + 16);
    for (int s = 0; s < V.S; ++s)
        dd_match_stream(g, V, s, det_tlwh, det_count, out_det_track_id, smem.data());
    double scratch[64];
    for (int s = 0; s < V.S; ++s)
        for (int d = 0; d < V.D; ++d) dd_apply_det(g, V, s, d, det_conf, det_label, scratch);
    return DD_OK;
}

int ddh_tracker_countline(void* state, const dd_tracker_config* cfg, const double* line,
                          int line_per_stream) {
    DDView V;
    int rc = dd_make_view(state, cfg, &V);
    if (rc != DD_OK) return rc;
    HostG g;
    for (int s = 0; s < V.S; ++s) dd_countline(g, V, s, line + (line_per_stream ? (size_t)s * 4 : 0));
    return DD_OK;
}

// scipy-exact LSAP on one dense f64 matrix (the device code path of dd_lsap).
int ddh_lsap(const double* cost, int nr, int nc, int32_t* out_col4row) {
    HostG g;
    const int n = nr > nc ? nr : nc;
    std::vector<char> mem(dd_lsap_scratch_bytes(n) + 16);
    DDLsapScratch s;
    dd_lsap_carve(mem.data(), n, s);
    struct Dense { const double* c; int nc; double operator()(int i, int j) const { return c[i * nc + j]; } };
    struct DenseT { const double* c; int nc; double operator()(int i, int j) const { return c[j * nc + i]; } };
    int rc;
    if (nc < nr) {
        DenseT f{cost, nc};
        rc = dd_lsap_solve(g, nc, nr, f, s);
        for (int r = 0; r < nr; ++r) out_col4row[r] = s.row4col[r];
    } else {
        Dense f{cost, nc};
        rc = dd_lsap_solve(g, nr, nc, f, s);
        for (int r = 0; r < nr; ++r) out_col4row[r] = s.col4row[r];
    }
    return rc;
}

int ddh_set_difference_order(const int32_t* a, int na, const int32_t* m, int nm, int32_t* out) {
    int maxv = 0;
    for (int i = 0; i < na; ++i) if (a[i] > maxv) maxv = a[i];
    for (int i = 0; i < nm; ++i) if (m[i] > maxv) maxv = m[i];
    std::vector<unsigned char> flag(maxv + 1, 0);
    for (int i = 0; i < nm; ++i) flag[m[i]] = 1;
    std::vector<short> av(na), o(na + 1);
    for (int i = 0; i < na; ++i) av[i] = (short)a[i];
    const int cap = dd_set_table_slots(na > 0 ? na : 1);
    std::vector<short> A(cap), B(cap), C(cap);
    bool contig = true;
    for (int i = 0; i < na; ++i) contig = contig && a[i] == i;
    int n;
    if (contig) {
        std::vector<short> surv;
        for (int i = 0; i < na; ++i) if (!flag[i]) surv.push_back((short)i);
        if ((na >> 2) > nm) { n = (int)surv.size(); for (int i = 0; i < n; ++i) o[i] = surv[i]; }
        else n = dd_set_order_from_survivors(surv.data(), (int)surv.size(), o.data(), A.data(), B.data());
    } else {
        n = dd_set_difference_order_serial(av.data(), na, flag.data(), nm, o.data(), A.data(), B.data(),
                                           C.data(), cap);
    }
    for (int i = 0; i < n; ++i) out[i] = o[i];
    return n;
}

int ddh_intersection(const double* seg, int n, int32_t* out) {
    for (int i = 0; i < n; ++i) {
        const double* p = seg + (size_t)i * 8;
        out[i] = dd_segments_intersect(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]) ? 1 : 0;
    }
    return 0;
}

}  // extern "C"
