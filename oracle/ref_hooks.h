/* oracle/ref_hooks.h — TEST INFRASTRUCTURE ONLY.
 * Hook table behind the OpenCV stand-in (oracle/shim/opencv2/core.hpp) used to compile the reference's own
 * FeatureExtractor.cpp / FeatureMatcher.cpp into oracle/_ref/.  All buffers are contiguous 8-bit single channel. */
#ifndef GVS_REF_HOOKS_H
#define GVS_REF_HOOKS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gvs_ref_hooks {
    /* cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR) */
    void (*resize_linear)(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
    /* cv::FAST(img, kps, threshold, true): returns n, fills pt.x, pt.y, response in OpenCV's output order */
    int (*fast_nms)(const uint8_t* img, int w, int h, int threshold, int* xs, int* ys, int* scores, int cap);
    /* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) */
    void (*gaussian_blur7)(const uint8_t* src, int w, int h, uint8_t* dst);
    /* cv::sortIdx(vector<float>, idx, SORT_DESCENDING) */
    void (*sort_idx_desc)(const float* key, int n, int* idx);
    /* cv::fastAtan2(y, x) */
    float (*fast_atan2)(float y, float x);
} gvs_ref_hooks;

/* current table (defaults: the restatements exported by liboracle.so) */
const gvs_ref_hooks* gvs_ref_get_hooks(void);
/* install another table (NULL restores the defaults); any NULL member keeps its default */
void gvs_ref_set_hooks(const gvs_ref_hooks* hooks);

#ifdef __cplusplus
}
#endif
#endif
