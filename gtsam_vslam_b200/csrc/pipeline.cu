// gvs_pipeline: a ring of independent slots (two extractors + one matcher each, own CUDA streams) so that a caller can keep
// several frames (or sub-batches) in flight: while the host collects the results of ticket t, tickets t+1 .. t+S-1 upload
// and compute.  SURVEY 8(f) rank 4.  The reference processes one frame at a time, synchronously, on the tracking thread
// (System::trackNewImage -> FeatureTracker::TrackImageT -> extractORBAndStereoMatch, src/FeatureTracker.cpp:1120-1145, :56-70);
// this is the asynchronous form of that call for callers that read frames ahead (src/VIOSlam.cpp:290-320 loops over a
// recorded sequence).  Built only from the public C ABI of include/gvs.h: no state of its own on the device.
#include <vector>

#include "../../include/gvs.h"
#include "common.cuh"

using namespace gvs;

struct gvs_pipeline {
    struct Slot {
        gvs_extractor* L = nullptr;
        gvs_extractor* R = nullptr;
        gvs_matcher* m = nullptr;
        int64_t ticket = -1;  // ticket currently owning the slot, -1 = free
        int n = 0;            // frames of that ticket
        bool done = false;    // wait() has returned for it
    };
    std::vector<Slot> slots;
    int width = 0, height = 0, framesPerSlot = 0;
    int64_t next = 0;
};

static gvs_pipeline::Slot* slotOf(gvs_pipeline* p, int64_t ticket) {
    if (!p || ticket < 0 || ticket >= p->next) return nullptr;
    gvs_pipeline::Slot& s = p->slots[(size_t)(ticket % (int64_t)p->slots.size())];
    return s.ticket == ticket ? &s : nullptr;
}

extern "C" {

int gvs_pipeline_create(gvs_pipeline** out, const gvs_extractor_params* params, int width, int height, int frames_per_slot,
                        int n_slots, double fx, float baseline, int device) {
    if (!out) return fail(GVS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_slots < 1 || n_slots > 64 || frames_per_slot < 1) return fail(GVS_ERR_INVALID, "bad slot count / frames per slot");
    gvs_pipeline* p = new gvs_pipeline();
    p->width = width; p->height = height; p->framesPerSlot = frames_per_slot;
    p->slots.resize((size_t)n_slots);
    for (auto& s : p->slots) {
        int rc = gvs_extractor_create(&s.L, params, width, height, frames_per_slot, device);
        if (!rc) rc = gvs_extractor_create(&s.R, params, width, height, frames_per_slot, device);
        if (!rc) rc = gvs_matcher_create(&s.m, s.L, s.R, fx, baseline, height);
        if (rc) { gvs_pipeline_destroy(p); return rc; }   // gvs_last_error() holds the cause
    }
    *out = p;
    return GVS_OK;
}

int gvs_pipeline_destroy(gvs_pipeline* p) {
    if (!p) return GVS_OK;
    for (auto& s : p->slots) {
        if (s.m) gvs_matcher_destroy(s.m);
        if (s.R) gvs_extractor_destroy(s.R);
        if (s.L) gvs_extractor_destroy(s.L);
    }
    delete p;
    return GVS_OK;
}

int gvs_pipeline_submit(gvs_pipeline* p, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride,
                        int n_frames, int64_t* ticket) {
    if (!p || !left || !right || !ticket) return fail(GVS_ERR_INVALID, "null argument");
    if (n_frames < 1 || n_frames > p->framesPerSlot) return fail(GVS_ERR_INVALID, "n_frames out of range for a slot");
    gvs_pipeline::Slot& s = p->slots[(size_t)(p->next % (int64_t)p->slots.size())];
    if (s.ticket >= 0) return fail(GVS_ERR_STATE, "pipeline full: release the oldest ticket first");
    int rc = gvs_track_stereo_batch_host(s.m, left, right, image_stride, row_stride, n_frames, p->width, p->height);
    if (rc) return rc;
    s.ticket = p->next++; s.n = n_frames; s.done = false;
    *ticket = s.ticket;
    return GVS_OK;
}

int gvs_pipeline_submit_temporal(gvs_pipeline* p, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride,
                                 int n_frames, const gvs_temporal_params* temporal, int64_t* ticket) {
    if (!temporal) return gvs_pipeline_submit(p, left, right, image_stride, row_stride, n_frames, ticket);
    if (!p || !left || !right || !ticket) return fail(GVS_ERR_INVALID, "null argument");
    if (n_frames < 1 || n_frames > p->framesPerSlot) return fail(GVS_ERR_INVALID, "n_frames out of range for a slot");
    gvs_pipeline::Slot& s = p->slots[(size_t)(p->next % (int64_t)p->slots.size())];
    if (s.ticket >= 0) return fail(GVS_ERR_STATE, "pipeline full: release the oldest ticket first");
    int rc = gvs_track_stereo_batch_host(s.m, left, right, image_stride, row_stride, n_frames, p->width, p->height);
    if (!rc) rc = gvs_track_temporal_batch(s.m, n_frames, temporal);
    if (rc) return rc;
    s.ticket = p->next++; s.n = n_frames; s.done = false;
    *ticket = s.ticket;
    return GVS_OK;
}

int gvs_pipeline_download_temporal(gvs_pipeline* p, int64_t ticket, int32_t* mp_offsets, int32_t* mp_key, int32_t* matches,
                                   int32_t* n_matches) {
    gvs_pipeline::Slot* s = slotOf(p, ticket);
    if (!s) return fail(GVS_ERR_INVALID, "unknown or released ticket");
    int rc = 0;
    if (!s->done && (rc = gvs_pipeline_wait(p, ticket, nullptr, nullptr))) return rc;
    return gvs_temporal_download(s->m, 0, s->n, mp_offsets, mp_key, matches, n_matches, nullptr, nullptr, nullptr, nullptr);
}

int gvs_pipeline_wait(gvs_pipeline* p, int64_t ticket, int32_t* left_offsets, int32_t* right_offsets) {
    gvs_pipeline::Slot* s = slotOf(p, ticket);
    if (!s) return fail(GVS_ERR_INVALID, "unknown or released ticket");
    int rc = gvs_matcher_sync(s->m);   // a device-side failure of this ticket (e.g. GVS_ERR_CAPACITY) surfaces here
    if (rc) return rc;
    s->done = true;
    if (left_offsets && (rc = gvs_extract_offsets(s->L, left_offsets))) return rc;
    if (right_offsets && (rc = gvs_extract_offsets(s->R, right_offsets))) return rc;
    return GVS_OK;
}

int gvs_pipeline_download(gvs_pipeline* p, int64_t ticket, gvs_keypoint* kp_left, uint8_t* desc_left, int cap_left,
                          gvs_keypoint* kp_right, uint8_t* desc_right, int cap_right, int32_t* right_idxs, int32_t* left_idxs,
                          float* estimated_depth, uint8_t* close) {
    gvs_pipeline::Slot* s = slotOf(p, ticket);
    if (!s) return fail(GVS_ERR_INVALID, "unknown or released ticket");
    int rc = 0;
    if (!s->done && (rc = gvs_pipeline_wait(p, ticket, nullptr, nullptr))) return rc;
    if ((kp_left || desc_left) && (rc = gvs_extract_download(s->L, 0, s->n, kp_left, desc_left, cap_left))) return rc;
    if ((kp_right || desc_right) && (rc = gvs_extract_download(s->R, 0, s->n, kp_right, desc_right, cap_right))) return rc;
    if (right_idxs || left_idxs || estimated_depth || close)
        rc = gvs_stereo_download(s->m, 0, s->n, right_idxs, left_idxs, estimated_depth, close);
    return rc;
}

int gvs_pipeline_handles(gvs_pipeline* p, int64_t ticket, gvs_extractor** fe_left, gvs_extractor** fe_right, gvs_matcher** fm) {
    gvs_pipeline::Slot* s = slotOf(p, ticket);
    if (!s) return fail(GVS_ERR_INVALID, "unknown or released ticket");
    if (fe_left) *fe_left = s->L;
    if (fe_right) *fe_right = s->R;
    if (fm) *fm = s->m;
    return GVS_OK;
}

int gvs_pipeline_release(gvs_pipeline* p, int64_t ticket) {
    gvs_pipeline::Slot* s = slotOf(p, ticket);
    if (!s) return fail(GVS_ERR_INVALID, "unknown or released ticket");
    s->ticket = -1; s->n = 0; s->done = false;   // the next submit on this slot is ordered after this ticket's work by the slot's streams
    return GVS_OK;
}

}  // extern "C"
