// dd_view.h -- typed view of the caller-owned tracker state blob (layout in include/deepdish_b200.h).
#pragma once
#include "../../include/deepdish_b200.h"
#include "dd_common.cuh"
#include "dd_lsap.cuh"

struct DDView {
    int S, T, D, B, C, DW;           // streams, slots, det capacity, budget, labels, gate words / row
    int tab_cap;                     // slots of each CPython-set emulation table (dd_set_table_slots(T))
    int max_age, n_init;
    double thr_cos, thr_iou;
    int lbl_motorbike, lbl_bicycle;
    int *n_tracks, *next_id, *n_deleted, *err, *order, *deleted;
    long long* counts;
    double *mean, *cov;
    int *track_id, *hits, *age, *tsu, *state, *gal_len, *gal_pos;
    float* gal;
    int* lab_cnt;
    double* lab_sum;
    int* path_n;
    double* path_last;
    int* path_crossed;
    unsigned* gate;
    float* cost;
    double* det_xyah;
    float* det_featn;
    int *det_slot, *det_kind;
    int* cdesc;
    int *work, *work_ctl;
    unsigned short *galh, *det_feath;
    int label_rank[DD_MAX_LABELS];
};

static inline uint64_t dd_align256(uint64_t x) { return (x + 255u) & ~(uint64_t)255u; }

static inline int dd_layout_compute(const dd_tracker_config* c, dd_tracker_layout* L) {
    if (!c || !L) return DD_ERR_INVALID;
    if (c->n_streams <= 0 || c->max_tracks <= 0 || c->max_dets <= 0 || c->budget <= 0) return DD_ERR_INVALID;
    if (c->feat_dim != DD_FEAT_DIM) return DD_ERR_INVALID;
    if (c->n_labels <= 0 || c->n_labels > DD_MAX_LABELS) return DD_ERR_INVALID;
    if (c->max_tracks > 1024 || c->max_dets > 1024) return DD_ERR_INVALID;
    if (c->budget > 32767) return DD_ERR_INVALID;      /* the track descriptor packs the gallery length in 15 bits */
    if (c->max_age < 0 || c->n_init < 1) return DD_ERR_INVALID;
    const uint64_t S = c->n_streams, T = c->max_tracks, D = c->max_dets, B = c->budget,
                   C = c->n_labels, F = DD_FEAT_DIM, DW = (D + 31) / 32;
    uint64_t off = 0;
#define DD_PUT(name, bytes) do { L->name = off; off = dd_align256(off + (uint64_t)(bytes)); } while (0)
    DD_PUT(n_tracks, 4 * S);
    DD_PUT(next_id, 4 * S);
    DD_PUT(n_deleted, 4 * S);
    DD_PUT(err, 4 * S);
    DD_PUT(order, 4 * S * T);
    DD_PUT(deleted, 4 * S * T);
    DD_PUT(counts, 8 * S * C * 4);
    DD_PUT(mean, 8 * S * T * 8);
    DD_PUT(cov, 8 * S * T * 64);
    DD_PUT(track_id, 4 * S * T);
    DD_PUT(hits, 4 * S * T);
    DD_PUT(age, 4 * S * T);
    DD_PUT(tsu, 4 * S * T);
    DD_PUT(state, 4 * S * T);
    DD_PUT(gal_len, 4 * S * T);
    DD_PUT(gal_pos, 4 * S * T);
    DD_PUT(gal, 4 * S * T * B * F);
    DD_PUT(galh, 2 * S * T * B * F);
    DD_PUT(lab_cnt, 4 * S * T * C);
    DD_PUT(lab_sum, 8 * S * T * C);
    DD_PUT(path_n, 4 * S * T);
    DD_PUT(path_last, 8 * S * T * 2);
    DD_PUT(path_crossed, 4 * S * T);
    DD_PUT(gate, 4 * S * T * DW);
    DD_PUT(cost, 4 * S * T * D);
    DD_PUT(det_xyah, 8 * S * D * 4);
    DD_PUT(det_featn, 4 * S * D * F);
    DD_PUT(det_slot, 4 * S * D);
    DD_PUT(det_kind, 4 * S * D);
    DD_PUT(cdesc, 4 * S * T * 2);
    DD_PUT(work, 4 * S * T);
    DD_PUT(work_ctl, 4 * 64);
    DD_PUT(det_feath, 2 * S * D * F);
#undef DD_PUT
    L->total_bytes = off;
    return DD_OK;
}

static inline int dd_make_view(void* blob, const dd_tracker_config* c, DDView* v) {
    dd_tracker_layout L;
    int rc = dd_layout_compute(c, &L);
    if (rc != DD_OK) return rc;
    if (!blob) return DD_ERR_INVALID;
    char* b = (char*)blob;
    v->S = c->n_streams; v->T = c->max_tracks; v->D = c->max_dets; v->B = c->budget;
    v->C = c->n_labels; v->DW = (c->max_dets + 31) / 32;
    v->tab_cap = dd_set_table_slots(c->max_tracks);
    v->max_age = c->max_age; v->n_init = c->n_init;
    v->thr_cos = c->max_cosine_distance; v->thr_iou = c->max_iou_distance;
    v->lbl_motorbike = c->label_motorbike; v->lbl_bicycle = c->label_bicycle;
    v->n_tracks = (int*)(b + L.n_tracks); v->next_id = (int*)(b + L.next_id);
    v->n_deleted = (int*)(b + L.n_deleted); v->err = (int*)(b + L.err);
    v->order = (int*)(b + L.order); v->deleted = (int*)(b + L.deleted);
    v->counts = (long long*)(b + L.counts);
    v->mean = (double*)(b + L.mean); v->cov = (double*)(b + L.cov);
    v->track_id = (int*)(b + L.track_id); v->hits = (int*)(b + L.hits); v->age = (int*)(b + L.age);
    v->tsu = (int*)(b + L.tsu); v->state = (int*)(b + L.state);
    v->gal_len = (int*)(b + L.gal_len); v->gal_pos = (int*)(b + L.gal_pos);
    v->gal = (float*)(b + L.gal);
    v->lab_cnt = (int*)(b + L.lab_cnt); v->lab_sum = (double*)(b + L.lab_sum);
    v->path_n = (int*)(b + L.path_n); v->path_last = (double*)(b + L.path_last);
    v->path_crossed = (int*)(b + L.path_crossed);
    v->gate = (unsigned*)(b + L.gate); v->cost = (float*)(b + L.cost);
    v->det_xyah = (double*)(b + L.det_xyah); v->det_featn = (float*)(b + L.det_featn);
    v->det_slot = (int*)(b + L.det_slot); v->det_kind = (int*)(b + L.det_kind);
    v->cdesc = (int*)(b + L.cdesc);
    v->galh = (unsigned short*)(b + L.galh); v->det_feath = (unsigned short*)(b + L.det_feath);
    v->work = (int*)(b + L.work); v->work_ctl = (int*)(b + L.work_ctl);
    for (int i = 0; i < DD_MAX_LABELS; ++i) v->label_rank[i] = i < c->n_labels ? c->label_rank[i] : 0;
    return DD_OK;
}
