// gvs_multi: one process, several GPUs.  Stereo frames are independent, so a batch of n frames is cut into
// contiguous ranges [g*n/G, (g+1)*n/G), one per device; every device has its own handle set (two extractors + one matcher, own
// streams), is driven by its own host thread and writes its results straight into the caller's dense output arrays (host gather:
// the only cross-device step is the prefix sum of the per-device key counts, done on the host).  No collective, no NCCL, no peer
// access on the data path.  The reference itself is single-GPU-less and single-frame (FeatureTracker::extractORBAndStereoMatch,
// src/FeatureTracker.cpp:56-70); this is that call for a batch of frames over several devices.  Built only from the public C ABI.
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/gvs.h"
#include "common.cuh"

using namespace gvs;

struct gvs_multi {
    struct Shard {
        int device = 0;
        gvs_extractor* L = nullptr;
        gvs_extractor* R = nullptr;
        gvs_matcher* m = nullptr;
        std::vector<int32_t> offL, offR, mpOff;   // per-frame offsets of the shard's last batch (local)
        int rc = 0;
        std::string err;
    };
    std::vector<Shard> shards;
    int width = 0, height = 0, framesPerDevice = 0;
};

namespace {

// a reusable barrier for the worker threads of one call
struct Barrier {
    std::mutex mu;
    std::condition_variable cv;
    int count, waiting = 0, generation = 0;
    explicit Barrier(int n) : count(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int gen = generation;
        if (++waiting == count) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != generation; });
    }
};

}  // namespace

extern "C" {

int gvs_multi_create(gvs_multi** out, const gvs_extractor_params* params, int width, int height, int frames_per_device, double fx,
                     float baseline, const int* devices, int n_devices) {
    if (!out || !params || !devices) return fail(GVS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || frames_per_device < 1) return fail(GVS_ERR_INVALID, "bad device count / frames per device");
    gvs_multi* p = new gvs_multi();
    p->width = width; p->height = height; p->framesPerDevice = frames_per_device;
    p->shards.resize((size_t)n_devices);
    for (int g = 0; g < n_devices; g++) {
        gvs_multi::Shard& s = p->shards[(size_t)g];
        s.device = devices[g];
        int rc = gvs_extractor_create(&s.L, params, width, height, frames_per_device, s.device);
        if (!rc) rc = gvs_extractor_create(&s.R, params, width, height, frames_per_device, s.device);
        if (!rc) rc = gvs_matcher_create(&s.m, s.L, s.R, fx, baseline, height);
        if (rc) { gvs_multi_destroy(p); return rc; }   // gvs_last_error() holds the cause
    }
    *out = p;
    return GVS_OK;
}

int gvs_multi_destroy(gvs_multi* p) {
    if (!p) return GVS_OK;
    for (auto& s : p->shards) {
        if (s.m) gvs_matcher_destroy(s.m);
        if (s.R) gvs_extractor_destroy(s.R);
        if (s.L) gvs_extractor_destroy(s.L);
    }
    delete p;
    return GVS_OK;
}

int gvs_multi_device_count(const gvs_multi* p) { return p ? (int)p->shards.size() : 0; }

int gvs_multi_track_stereo(gvs_multi* p, const uint8_t* left, const uint8_t* right, size_t image_stride, int row_stride, int n_frames,
                           const gvs_temporal_params* temporal, int32_t* left_offsets, int32_t* right_offsets, gvs_keypoint* kp_left,
                           uint8_t* desc_left, int cap_left, gvs_keypoint* kp_right, uint8_t* desc_right, int cap_right,
                           int32_t* right_idxs, int32_t* left_idxs, float* estimated_depth, uint8_t* close, int32_t* mp_offsets,
                           int32_t* mp_key, int32_t* matches, int cap_mp, int32_t* n_matches) {
    if (!p || !left || !right || !left_offsets || !right_offsets) return fail(GVS_ERR_INVALID, "null argument");
    const int G = (int)p->shards.size();
    if (n_frames < 1 || n_frames > G * p->framesPerDevice) return fail(GVS_ERR_INVALID, "n_frames out of range for these devices");
    if (temporal && (!temporal->prev || !temporal->pose_prev || !temporal->to_camera || !temporal->to_camera_r || !mp_offsets))
        return fail(GVS_ERR_INVALID, "temporal parameter arrays missing");
    // contiguous frame ranges; with temporal tracking a frame and its predecessor must sit on the same device
    std::vector<int> first((size_t)G + 1);
    for (int g = 0; g <= G; g++) first[(size_t)g] = (int)((long long)g * n_frames / G);
    if (temporal)
        for (int g = 0; g < G; g++)
            for (int f = first[(size_t)g]; f < first[(size_t)g + 1]; f++) {
                const int q = temporal->prev[f];
                if (q >= 0 && (q < first[(size_t)g] || q >= first[(size_t)g + 1]))
                    return fail(GVS_ERR_INVALID, "prev[] crosses a device boundary: cut the batch where the sequences are cut");
            }
    Barrier barrier(G);
    std::vector<int> baseL((size_t)G + 1, 0), baseR((size_t)G + 1, 0), baseM((size_t)G + 1, 0);
    std::atomic<int> failed{0};
    auto work = [&](int g) {
        gvs_multi::Shard& s = p->shards[(size_t)g];
        const int f0 = first[(size_t)g], nf = first[(size_t)g + 1] - f0;
        s.rc = 0; s.err.clear();
        auto check = [&](int rc) { if (rc && !s.rc) { s.rc = rc; s.err = gvs_last_error(); failed = 1; } return rc; };
        s.offL.assign((size_t)nf + 1, 0); s.offR.assign((size_t)nf + 1, 0); s.mpOff.assign((size_t)nf + 1, 0);
        if (nf > 0) {
            if (!check(gvs_track_stereo_batch_host(s.m, left + (size_t)f0 * image_stride, right + (size_t)f0 * image_stride, image_stride,
                                                   row_stride, nf, p->width, p->height)) && temporal) {
                gvs_temporal_params t = *temporal;
                std::vector<int32_t> prev((size_t)nf);
                for (int f = 0; f < nf; f++) prev[(size_t)f] = temporal->prev[f0 + f] < 0 ? -1 : temporal->prev[f0 + f] - f0;
                t.prev = prev.data(); t.pose_prev = temporal->pose_prev + 16 * (size_t)f0; t.to_camera = temporal->to_camera + 16 * (size_t)f0;
                t.to_camera_r = temporal->to_camera_r + 16 * (size_t)f0;
                check(gvs_track_temporal_batch(s.m, nf, &t));
            }
            if (!s.rc) check(gvs_extract_offsets(s.L, s.offL.data()));
            if (!s.rc) check(gvs_extract_offsets(s.R, s.offR.data()));
            if (!s.rc && temporal) check(gvs_temporal_download(s.m, 0, nf, s.mpOff.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        }
        barrier.wait();                       // every shard knows its counts
        if (g == 0) {
            for (int k = 0; k < G; k++) {
                baseL[(size_t)k + 1] = baseL[(size_t)k] + p->shards[(size_t)k].offL.back();
                baseR[(size_t)k + 1] = baseR[(size_t)k] + p->shards[(size_t)k].offR.back();
                baseM[(size_t)k + 1] = baseM[(size_t)k] + p->shards[(size_t)k].mpOff.back();
            }
            if (!failed && (baseL[(size_t)G] > cap_left || baseR[(size_t)G] > cap_right || (temporal && (mp_key || matches) && baseM[(size_t)G] > cap_mp))) {
                s.rc = fail(GVS_ERR_CAPACITY, "caller buffers too small for the keys / map points of this batch");
                s.err = gvs_last_error(); failed = 1;
            }
        }
        barrier.wait();                       // the global offsets are known
        if (failed || nf == 0) return;
        const int bl = baseL[(size_t)g], br = baseR[(size_t)g], bm = baseM[(size_t)g];
        for (int f = 0; f < nf; f++) { left_offsets[f0 + f] = bl + s.offL[(size_t)f]; right_offsets[f0 + f] = br + s.offR[(size_t)f]; }
        if (g == G - 1) { left_offsets[n_frames] = baseL[(size_t)G]; right_offsets[n_frames] = baseR[(size_t)G]; }
        check(gvs_extract_download(s.L, 0, nf, kp_left ? kp_left + bl : nullptr, desc_left ? desc_left + (size_t)bl * 32 : nullptr, cap_left - bl));
        check(gvs_extract_download(s.R, 0, nf, kp_right ? kp_right + br : nullptr, desc_right ? desc_right + (size_t)br * 32 : nullptr, cap_right - br));
        if (right_idxs || left_idxs || estimated_depth || close)
            check(gvs_stereo_download(s.m, 0, nf, right_idxs ? right_idxs + bl : nullptr, left_idxs ? left_idxs + br : nullptr,
                                      estimated_depth ? estimated_depth + bl : nullptr, close ? close + bl : nullptr));
        if (temporal) {
            for (int f = 0; f < nf; f++) mp_offsets[f0 + f] = bm + s.mpOff[(size_t)f];
            if (g == G - 1) mp_offsets[n_frames] = baseM[(size_t)G];
            check(gvs_temporal_download(s.m, 0, nf, nullptr, mp_key ? mp_key + bm : nullptr, matches ? matches + 2 * (size_t)bm : nullptr,
                                        n_matches ? n_matches + f0 : nullptr, nullptr, nullptr, nullptr, nullptr));
        }
    };
    std::vector<std::thread> threads;
    for (int g = 1; g < G; g++) threads.emplace_back(work, g);
    work(0);
    for (auto& t : threads) t.join();
    for (auto& s : p->shards)
        if (s.rc) return fail(s.rc, "gvs_multi_track_stereo", s.err.c_str());
    return GVS_OK;
}

}  // extern "C"
