// TEST INFRASTRUCTURE ONLY -- CPU oracle #2 and CPU baseline: a scalar C++17 restatement of the
// reference's frontend hot path with NO OpenCV dependency.
//
// Nothing under gtsam_vslam_b200/ links, loads or calls this file.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs do.
//
// Parity status: PINNED against outputs of the reference itself run here.  The reference's own
// FeatureExtractor.cpp / FeatureMatcher.cpp are compiled unmodified into oracle/_ref (oracle/Makefile target
// `ref`, against the stand-ins of oracle/shim) and, with the real OpenCV 4.13.0 primitives hooked in through
// cv2, reproduce this file bit for bit (tests/test_reference_build.py).  It is also pinned against
// (a) oracle/ref_cv2.py, which runs the same control flow on the real OpenCV primitives, primitive by
// primitive and end to end (tests/test_oracle_vs_cv2.py), and (b) the committed fixtures tests/golden/*.npz.
// The reference holds no tests or golden vectors of its own.
//
// Third-party dependency restated here: OpenCV (requested as 4.2.0 else 3.2, reference
// CMakeLists.txt:27-33, un-vendored).  Restated primitives: cv::resize INTER_LINEAR u8, cv::FAST 9/16
// + NMS, cv::sortIdx (libstdc++ std::sort + reverse), cv::fastAtan2, cv::GaussianBlur 7x7 sigma 2 u8
// fixed point, cv::norm L1, cvRound/cvFloor/cvCeil.  Semantics: SURVEY.md Appendix A.
//
// Build: see oracle/Makefile (-O3 -march=native as reference CMakeLists.txt:23-26, plus -ffp-contract=off
// so that float32 expressions round exactly like the un-fused canonical definition).
//
// All citations are file:line under /root/reference.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <thread>
#include <utility>
#include <vector>

namespace {

const signed char kPattern[1024] = {
#include "../include/gvs_orb_pattern.inc"
};

inline int cvRoundF(float v) { return (int)lrintf(v); }    // half to even (default rounding mode)
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { return (int)std::floor(v); }
inline int cvCeilF(float v) { return (int)std::ceil(v); }

struct KeyPoint {  // cv::KeyPoint layout, 28 bytes
    float x, y, size, angle, response;
    int octave, class_id;
};

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    const uint8_t* row(int y) const { return d.data() + (size_t)y * w; }
    uint8_t* row(int y) { return d.data() + (size_t)y * w; }
};

// ---------------------------------------------------------------- cv::resize INTER_LINEAR, 8UC1 (App. A.1)
void resizeLinear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    const int ONE = 2048;
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    int xmax = dw;
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorF(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { xmax = std::min(xmax, dx); fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cvRoundF((1.f - fx) * ONE);
        alpha[2 * dx + 1] = (short)cvRoundF(fx * ONE);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorF(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cvRoundF((1.f - fy) * ONE);
        beta[2 * dy + 1] = (short)cvRoundF(fy * ONE);
    }
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int>& out) {
        sy = std::min(std::max(sy, 0), sh - 1);
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            out[dx] = dx < xmax ? S[sx] * alpha[2 * dx] + S[sx + 1] * alpha[2 * dx + 1] : S[sx] * ONE;
        }
    };
    for (int dy = 0; dy < dh; dy++) {
        hrow(yofs[dy], r0);
        hrow(yofs[dy] + 1, r1);
        int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int x = 0; x < dw; x++)
            D[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------- cv::GaussianBlur 7x7 sigma=2, 8U fixed point (App. A.2)
inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}
void gaussianBlur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    static const int T[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* H = hbuf.data() + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3) for (int k = 0; k < 7; k++) acc += T[k] * S[x + k - 3];
            else for (int k = 0; k < 7; k++) acc += T[k] * S[reflect101(x + k - 3, w)];
            H[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; k++) R[k] = hbuf.data() + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)T[k] * R[k][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------- cv::fastAtan2 (App. A.4)
float fastAtan2(float y, float x) {
    const float s = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s,
                p7 = -0.04432655554792128f * s;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------- cv::FAST TYPE_9_16 with NMS on a sub-image (App. A.3)
struct FastPt { int x, y, score; };
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// max over the 16 arcs of 9 consecutive ring pixels of max(min d, min -d), d = centre - ring
inline int fastBest(const uint8_t* p, const int* off) {
    int d[25];
    int v = p[0];
    for (int i = 0; i < 16; i++) d[i] = v - p[off[i]];
    for (int i = 16; i < 25; i++) d[i] = d[i - 16];
    int best = 0;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

void fastDetect(const uint8_t* img, int stride, int w, int h, int threshold, std::vector<FastPt>& out,
                std::vector<uint8_t>& score /*scratch w*h*/) {
    out.clear();
    if (w < 7 || h < 7) return;
    int off[16];
    for (int i = 0; i < 16; i++) off[i] = kRingDy[i] * stride + kRingDx[i];
    score.assign((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t* r = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = r + x;
            int v = p[0];
            // any 9-arc contains ring pixel k or k+8 for every k: cheap rejects first
            int a = v - p[off[0]], b = v - p[off[8]];
            if (std::abs(a) <= threshold && std::abs(b) <= threshold) continue;
            a = v - p[off[4]]; b = v - p[off[12]];
            if (std::abs(a) <= threshold && std::abs(b) <= threshold) continue;
            int best = fastBest(p, off);
            if (best > threshold) score[(size_t)y * w + x] = (uint8_t)(best - 1);
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = score[(size_t)y * w + x];
            if (!s) continue;
            const uint8_t* c = &score[(size_t)y * w + x];
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] && s > c[w] &&
                s > c[w + 1])
                out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------- cv::sortIdx(row, SORT_DESCENDING) (App. A.6)
void sortIdxDescending(const float* key, int n, int* idx) {
    std::iota(idx, idx + n, 0);
    std::sort(idx, idx + n, [key](int a, int b) { return key[a] < key[b]; });
    for (int j = 0; j < n / 2; j++) std::swap(idx[j], idx[n - 1 - j]);
}

// ---------------------------------------------------------------- FeatureExtractor
struct LevelKeys {
    std::vector<int> x, y;
    std::vector<float> response, angle;
    std::vector<int> m10, m01;
};

struct Extractor {
    int nFeatures, nLevels, edgeThreshold, patchSize, maxFast, minFast;
    float imScale;
    const int halfPatchSize = 15;
    std::vector<float> scalePyramid, scaleInvPyramid, sigmaFactor, InvSigmaFactor;
    std::vector<int> scaledPatchSize, featurePerLevel, umax;
    std::vector<Image> imagePyramid, blurred;
    std::vector<LevelKeys> cands, kept;  // per level: pre-SSC candidates, post-SSC keys (+angle)
    std::vector<int> sscPasses;

    // src/FeatureExtractor.cpp:620-682
    Extractor(int nF, int nL, float s, int edge, int patch, int maxF, int minF)
        : nFeatures(nF), nLevels(nL), edgeThreshold(edge), patchSize(patch), maxFast(maxF), minFast(minF), imScale(s) {
        scalePyramid.assign(nL, 1.f); scaleInvPyramid.assign(nL, 1.f); sigmaFactor.assign(nL, 1.f);
        InvSigmaFactor.assign(nL, 1.f); scaledPatchSize.assign(nL, patch);
        for (int i = 1; i < nL; i++) {
            scalePyramid[i] = scalePyramid[i - 1] * imScale;
            scaledPatchSize[i] = (int)(patchSize * scalePyramid[i]);
            sigmaFactor[i] = scalePyramid[i] * scalePyramid[i];
        }
        for (int i = 0; i < nL; i++) { scaleInvPyramid[i] = 1.0f / scalePyramid[i]; InvSigmaFactor[i] = 1.0f / sigmaFactor[i]; }
        featurePerLevel.assign(nL, 0);
        float factor = 1.0f / imScale;
        float nd = nFeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nL));
        int sum = 0;
        for (int l = 0; l < nL - 1; l++) { featurePerLevel[l] = cvRoundF(nd); sum += featurePerLevel[l]; nd *= factor; }
        featurePerLevel[nL - 1] = std::max(nFeatures - sum, 0);
        umax.assign(halfPatchSize + 1, 0);
        int v, v0, vmax = cvFloorF(halfPatchSize * std::sqrt(2.f) / 2 + 1);
        int vmin = cvCeilF(halfPatchSize * std::sqrt(2.f) / 2);
        const double hp2 = halfPatchSize * halfPatchSize;
        for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(std::sqrt(hp2 - v * v));
        for (v = halfPatchSize, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // src/FeatureExtractor.cpp:342-366 (border frame is dead data)
    void computePyramid(const uint8_t* img, int w, int h, int stride) {
        imagePyramid.resize(nLevels);
        for (int l = 0; l < nLevels; l++) {
            float scale = scaleInvPyramid[l];
            int lw = cvRoundF((float)w * scale), lh = cvRoundF((float)h * scale);
            Image& im = imagePyramid[l];
            im.w = lw; im.h = lh; im.d.resize((size_t)lw * lh);
            if (l == 0) for (int y = 0; y < h; y++) std::memcpy(im.row(y), img + (size_t)y * stride, w);
            else {
                const Image& p = imagePyramid[l - 1];
                resizeLinear(p.d.data(), p.w, p.h, p.w, im.d.data(), lw, lh, lw);
            }
        }
    }

    // src/FeatureExtractor.cpp:315-340
    float computeOrientation(const Image& im, int px, int py, int* m10o, int* m01o) const {
        int m_01 = 0, m_10 = 0;
        const uint8_t* center = im.row(py) + px;
        for (int u = -halfPatchSize; u <= halfPatchSize; ++u) m_10 += u * center[u];
        int step = im.w;
        for (int v = 1; v <= halfPatchSize; ++v) {
            int v_sum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        if (m10o) *m10o = m_10;
        if (m01o) *m01o = m_01;
        return fastAtan2((float)m_01, (float)m_10);
    }

    // src/FeatureExtractor.cpp:368-468.  Returns indices (into the input order) of the kept points, in output order.
    std::vector<int> ssc(const LevelKeys& k, int numRetPoints, float tolerance, int cols, int rows, int* passesOut) const {
        int n = (int)k.x.size();
        std::vector<int> Indx(n);
        sortIdxDescending(k.response.data(), n, Indx.data());
        int exp1 = rows + cols + 2 * numRetPoints;
        long long exp2 = ((long long)4 * cols + (long long)4 * numRetPoints + (long long)4 * rows * numRetPoints +
                          (long long)rows * rows + (long long)cols * cols - (long long)2 * rows * cols +
                          (long long)4 * rows * cols * numRetPoints);
        double exp3 = std::sqrt((double)exp2);
        double exp4 = numRetPoints - 1;
        double sol1 = -std::round((exp1 + exp3) / exp4);
        double sol2 = -std::round((exp1 - exp3) / exp4);
        int high = (int)((sol1 > sol2) ? sol1 : sol2);
        int low = (int)std::floor(std::sqrt((double)n / numRetPoints));
        low = std::max(1, low);
        int width, prevWidth = -1, passes = 0;
        std::vector<int> ResultVec, result;
        unsigned K = numRetPoints;
        unsigned Kmin = (unsigned)std::round(K - (K * tolerance));
        unsigned Kmax = (unsigned)std::round(K + (K * tolerance));
        result.reserve(n);
        while (true) {
            width = low + (high - low) / 2;
            if (width == prevWidth || low > high) { ResultVec = result; break; }
            result.clear();
            double c = (double)width / 2.0;
            int numCellCols = (int)std::floor(cols / c), numCellRows = (int)std::floor(rows / c);
            std::vector<std::vector<bool>> covered(numCellRows + 1, std::vector<bool>(numCellCols + 1, false));
            passes++;
            for (int i = 0; i < n; ++i) {
                int row = (int)std::floor(k.y[Indx[i]] / c), col = (int)std::floor(k.x[Indx[i]] / c);
                if (!covered[row][col]) {
                    result.push_back(i);
                    int reach = (int)std::floor(width / c);
                    int rowMin = std::max(row - reach, 0), rowMax = std::min(row + reach, numCellRows);
                    int colMin = std::max(col - reach, 0), colMax = std::min(col + reach, numCellCols);
                    for (int r = rowMin; r <= rowMax; ++r)
                        for (int cc = colMin; cc <= colMax; ++cc) covered[r][cc] = true;
                }
            }
            if (result.size() >= Kmin && result.size() <= Kmax) { ResultVec = result; break; }
            else if (result.size() < Kmin) high = width - 1;
            else low = width + 1;
            prevWidth = width;
        }
        if (passesOut) *passesOut = passes;
        std::vector<int> out(ResultVec.size());
        for (size_t i = 0; i < ResultVec.size(); i++) out[i] = Indx[ResultVec[i]];
        return out;
    }

    // src/FeatureExtractor.cpp:535-618
    void computeKeypoints() {
        const int fastEdge = 3;
        const float W = 35;
        const int minX = edgeThreshold - fastEdge, minY = edgeThreshold - fastEdge;
        cands.assign(nLevels, LevelKeys());
        kept.assign(nLevels, LevelKeys());
        sscPasses.assign(nLevels, 0);
        std::vector<FastPt> temp;
        std::vector<uint8_t> scratch;
        for (int level = 0; level < nLevels; level++) {
            const Image& im = imagePyramid[level];
            const int maxX = im.w - edgeThreshold + fastEdge, maxY = im.h - edgeThreshold + fastEdge;
            const int wid = maxX - minX, hig = maxY - minY;
            const int nCols = (int)(wid / W), nRows = (int)(hig / W);
            LevelKeys& ck = cands[level];
            if (nCols > 0 && nRows > 0) {
                const int gridW = cvCeilF((float)wid / nCols), gridH = cvCeilF((float)hig / nRows);
                for (int iR = 0; iR < nRows; iR++) {
                    const int rStart = minY + iR * gridH;
                    int rEnd = rStart + gridH + 2 * fastEdge;
                    if (rStart >= maxY - 2 * fastEdge) continue;
                    if (rEnd > maxY) rEnd = maxY;
                    for (int iC = 0; iC < nCols; iC++) {
                        const int cStart = minX + iC * gridW;
                        int cEnd = cStart + gridW + 2 * fastEdge;
                        if (cStart >= maxX - 2 * fastEdge) continue;
                        if (cEnd > maxX) cEnd = maxX;
                        const uint8_t* sub = im.row(rStart) + cStart;
                        fastDetect(sub, im.w, cEnd - cStart, rEnd - rStart, maxFast, temp, scratch);
                        if (temp.empty()) fastDetect(sub, im.w, cEnd - cStart, rEnd - rStart, minFast, temp, scratch);
                        for (const FastPt& p : temp) {
                            ck.x.push_back(p.x + cStart);
                            ck.y.push_back(p.y + rStart);
                            ck.response.push_back((float)p.score);
                        }
                    }
                }
            }
            LevelKeys& kk = kept[level];
            const int fPLevel = featurePerLevel[level];
            if ((int)ck.x.size() > fPLevel) {
                std::vector<int> keep = ssc(ck, fPLevel, 0.1f, im.w, im.h, &sscPasses[level]);
                for (int i : keep) { kk.x.push_back(ck.x[i]); kk.y.push_back(ck.y[i]); kk.response.push_back(ck.response[i]); }
            } else {
                kk.x = ck.x; kk.y = ck.y; kk.response = ck.response;
            }
            size_t n = kk.x.size();
            kk.angle.resize(n); kk.m10.resize(n); kk.m01.resize(n);
            for (size_t i = 0; i < n; i++) kk.angle[i] = computeOrientation(im, kk.x[i], kk.y[i], &kk.m10[i], &kk.m01[i]);
        }
    }

    // src/FeatureExtractor.cpp:268-305
    void orbDescriptor(const Image& img, int cx, int cy, float kptAngle, uint8_t* desc) const {
        const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
        float angle = kptAngle * factorPI;
        float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
        const uint8_t* center = img.row(cy) + cx;
        const int step = img.w;
        const signed char* pat = kPattern;
        auto get = [&](int idx) -> int {
            float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
            return center[cvRoundF(px * b + py * a) * step + cvRoundF(px * a - py * b)];
        };
        for (int i = 0; i < 32; ++i, pat += 32) {
            int val = 0;
            for (int j = 0; j < 8; j++) {
                int t0 = get(2 * j), t1 = get(2 * j + 1);
                val |= (t0 < t1) << j;
            }
            desc[i] = (uint8_t)val;
        }
    }

    // src/FeatureExtractor.cpp:481-533
    int extract(const uint8_t* img, int w, int h, int stride, std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc) {
        computePyramid(img, w, h, stride);
        computeKeypoints();
        int nF = 0;
        for (int l = 0; l < nLevels; l++) nF += (int)kept[l].x.size();
        blurred.assign(nLevels, Image());
        if (nF <= 0) return 0;  // outputs untouched (:498-499)
        kps.assign(nF, KeyPoint());
        desc.assign((size_t)nF * 32, 0);
        int o = 0;
        for (int l = 0; l < nLevels; l++) {
            const LevelKeys& kk = kept[l];
            if (kk.x.empty()) continue;
            const float scale = scalePyramid[l];
            const Image& src = imagePyramid[l];
            Image& im = blurred[l];
            im.w = src.w; im.h = src.h; im.d.resize(src.d.size());
            gaussianBlur7(src.d.data(), src.w, src.h, src.w, im.d.data(), im.w);
            for (size_t i = 0; i < kk.x.size(); i++, o++) {
                orbDescriptor(im, kk.x[i], kk.y[i], kk.angle[i], &desc[(size_t)o * 32]);
                KeyPoint& kp = kps[o];
                kp.x = (float)kk.x[i]; kp.y = (float)kk.y[i];
                if (l != 0) { kp.x *= scale; kp.y *= scale; }
                kp.size = (float)scaledPatchSize[l];
                kp.angle = kk.angle[i];
                kp.response = kk.response[i];
                kp.octave = l;
                kp.class_id = -1;
            }
        }
        return nF;
    }
};

// ---------------------------------------------------------------- FeatureMatcher
// src/FeatureMatcher.cpp:710-726
inline int descriptorDistance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        unsigned v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

inline int sadL1(const Image& a, int ax, int ay, const Image& b, int bx, int by, int n) {
    int s = 0;
    for (int r = 0; r < n; r++) {
        const uint8_t* pa = a.row(ay + r) + ax; const uint8_t* pb = b.row(by + r) + bx;
        for (int c = 0; c < n; c++) s += std::abs((int)pa[c] - (int)pb[c]);
    }
    return s;
}

// src/FeatureMatcher.cpp:528-708 (+ destributeRightKeys :728-752)
void stereoMatch(const Extractor& feL, const Extractor& feR, double fx, float mBaseline, int imageHeight,
                 const KeyPoint* kl, const uint8_t* dl, int NL, const KeyPoint* kr, const uint8_t* dr, int NR,
                 int* rightIdxs, int* leftIdxs, float* estimatedDepth, uint8_t* close, int* statsOut) {
    const int thDist = 75, closeNumber = 40;
    std::vector<std::vector<int>> indexes(imageHeight);
    for (int i = 0; i < NR; i++) {
        const int yKey = cvRoundF(kr[i].y);
        const float r = 2.0f * feR.scalePyramid[kr[i].octave];
        const int mn = cvFloorF(yKey - r), mx = cvCeilF(yKey + r);
        for (int pos = mn; pos <= mx; pos++) indexes[pos].push_back(i);
    }
    for (int i = 0; i < NL; i++) { estimatedDepth[i] = -1.0f; close[i] = 0; rightIdxs[i] = -1; }
    for (int i = 0; i < NR; i++) leftIdxs[i] = -1;
    const float minD = 0, maxD = (float)fx;
    const int windowRadius = 5, windowMovementX = 5;
    std::vector<std::pair<float, int>> allDepths;
    std::vector<std::pair<int, int>> allDists2;
    int nHam = 0;
    for (int leftRow = 0; leftRow < NL; leftRow++) {
        const KeyPoint& it = kl[leftRow];
        const int yKey = cvRoundF(it.y);
        const float uL = it.y;
        const int octL = it.octave;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 256, bestIdx = -1;
        const std::vector<int>& bucket = indexes[yKey];
        if (bucket.empty()) continue;
        for (int idx : bucket) {
            const float uR = kr[idx].y;
            const int octR = kr[idx].octave;
            if (octR < octL - 1 || octR > octL + 1) continue;
            if (!(uR >= minU && uR <= maxU)) continue;
            int dist = descriptorDistance(dl + (size_t)leftRow * 32, dr + (size_t)idx * 32);
            nHam++;
            if (bestDist > dist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist > thDist) continue;
        const float kRx = kr[bestIdx].x;
        const float scale = feL.scaleInvPyramid[octL];
        const float scuL = std::round(it.x * scale), scvL = std::round(it.y * scale), scuR = std::round(kRx * scale);
        const Image& imL = feL.imagePyramid[octL];
        const Image& imR = feR.imagePyramid[octL];
        int bestDistW = INT_MAX, bestX = 0;
        float allDists[2 * windowMovementX + 1] = {0};
        for (int xMov = -windowMovementX; xMov <= windowMovementX; xMov++) {
            const float startW = scuR + xMov - windowRadius, endW = scuR + xMov + windowRadius + 1;
            if (startW < 0 || endW >= imR.w) continue;
            float dist = (float)sadL1(imL, (int)(scuL - windowRadius), (int)(scvL - windowRadius), imR, (int)startW,
                                      (int)(scvL - windowRadius), 2 * windowRadius + 1);
            if ((float)bestDistW > dist) { bestX = xMov; bestDistW = (int)dist; }
            allDists[xMov + windowMovementX] = dist;
        }
        if (bestX == -windowMovementX || bestX == windowMovementX) continue;
        const float dist1 = allDists[windowMovementX + bestX - 1], dist2 = allDists[windowMovementX + bestX],
                    dist3 = allDists[windowMovementX + bestX + 1];
        const float delta = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (delta > 1 || delta < -1) continue;
        const float newuR = feL.scalePyramid[octL] * ((float)scuR + (float)bestX + delta);
        const float disparity = it.x - newuR;
        if (disparity > 0.0f && disparity < fx) {
            const float depth = ((float)fx * mBaseline) / disparity;
            rightIdxs[leftRow] = bestIdx;
            leftIdxs[bestIdx] = leftRow;
            estimatedDepth[leftRow] = depth;
            allDists2.emplace_back(bestDistW, leftRow);
            allDepths.emplace_back(depth, leftRow);
            if (depth < mBaseline * closeNumber) close[leftRow] = 1;
        }
    }
    if (statsOut) { statsOut[0] = nHam; statsOut[1] = (int)allDepths.size(); }
    if (allDepths.empty()) return;
    std::sort(allDepths.begin(), allDepths.end());
    std::sort(allDists2.begin(), allDists2.end());
    const int medianD = allDists2[allDists2.size() / 2].first;
    const float medDistD = (float)medianD * (1.5f * 1.4f);
    const int endDe = (int)std::floor(allDepths.size() * 0.01);
    auto drop = [&](int li) {
        const int rIdx = rightIdxs[li];
        if (rIdx >= 0) leftIdxs[rIdx] = -1;
        rightIdxs[li] = -1; estimatedDepth[li] = -1; close[li] = 0;
    };
    for (int i = 0; i < endDe; i++) drop(allDepths[i].second);
    for (int i = (int)allDists2.size() - 1; i >= 0; i--) {
        if (allDists2[i].first < medDistD) break;
        drop(allDists2[i].second);
    }
}

// src/FeatureTracker.cpp:28-54.  CSR form of the vector<vector<vector<int>>> grid (cell-major, insertion order).
struct Grid {
    int xGrids = 0, yGrids = 0;
    float xMult = 0, yMult = 0;
    std::vector<int> start, items;  // start[yGrids*xGrids+1]
};
Grid assignKeysToGrids(const KeyPoint* k, int n, int width, int height) {
    Grid g;
    const float imageRatio = (float)width / (float)height;
    g.xGrids = 64;
    g.yGrids = cvCeilF((float)g.xGrids / imageRatio);
    g.xMult = (float)g.xGrids / (float)width;
    g.yMult = (float)g.yGrids / (float)height;
    std::vector<std::vector<int>> cells((size_t)g.xGrids * g.yGrids);
    for (int i = 0; i < n; i++) {
        int xPos = cvRoundF(k[i].x * g.xMult), yPos = cvRoundF(k[i].y * g.yMult);
        xPos = std::min(std::max(xPos, 0), g.xGrids - 1);
        yPos = std::min(std::max(yPos, 0), g.yGrids - 1);
        cells[(size_t)yPos * g.xGrids + xPos].push_back(i);
    }
    g.start.assign(cells.size() + 1, 0);
    for (size_t c = 0; c < cells.size(); c++) {
        g.start[c + 1] = g.start[c] + (int)cells[c].size();
        g.items.insert(g.items.end(), cells[c].begin(), cells[c].end());
    }
    return g;
}

// src/FeatureMatcher.cpp:13-64
void getMatchIdxs(float px, float py, std::vector<int>& idxs, const Grid& g, const KeyPoint* kps, int predictedScale,
                  float radius) {
    const float trackX = px, trackY = py;
    const int minX = std::max(0, cvFloorF((trackX - radius) * g.xMult));
    const int maxX = std::min(g.xGrids - 1, cvCeilF((trackX + radius) * g.xMult));
    const int minY = std::max(0, cvFloorF((trackY - radius) * g.yMult));
    const int maxY = std::min(g.yGrids - 1, cvCeilF((trackY + radius) * g.yMult));
    if (minX >= g.xGrids || minY >= g.yGrids || maxX < 0 || maxY < 0) return;
    const int maxLevel = predictedScale + 1, minLevel = predictedScale - 1;
    for (int row = minY; row <= maxY; row++)
        for (int col = minX; col <= maxX; col++) {
            int c = row * g.xGrids + col;
            for (int j = g.start[c]; j < g.start[c + 1]; j++) {
                const KeyPoint& kp = kps[g.items[j]];
                if (kp.octave > maxLevel || kp.octave < minLevel) continue;
                const float distx = kp.x - trackX, disty = kp.y - trackY;
                if (std::fabs(distx) < radius && std::fabs(disty) < radius) idxs.push_back(g.items[j]);
            }
        }
}

struct Best { int bestDist = 256, bestIdx = -1, bestLev = -1, bestLev2 = -1, secDist = 256; };

template <class Skip>
Best scanCandidates(const uint8_t* mpDesc, const std::vector<int>& idxs, const int* matched, const KeyPoint* kps,
                    const uint8_t* descs, Skip skip) {
    Best b;
    for (int idx : idxs) {
        if (matched[idx] >= 0) continue;
        if (skip(idx)) continue;
        const int lev = kps[idx].octave;
        int dist = descriptorDistance(mpDesc, descs + (size_t)idx * 32);
        if (dist < b.bestDist) { b.secDist = b.bestDist; b.bestLev2 = b.bestLev; b.bestDist = dist; b.bestLev = lev; b.bestIdx = idx; continue; }
        if (dist < b.secDist) { b.secDist = dist; b.bestLev2 = lev; }
    }
    return b;
}

}  // namespace

// ================================================================ C API (ctypes; tests and bench baseline only)
extern "C" {

void* ora_extractor_create(int nF, int nL, float s, int edge, int patch, int maxF, int minF) {
    return new Extractor(nF, nL, s, edge, patch, maxF, minF);
}
void ora_extractor_destroy(void* h) { delete (Extractor*)h; }

void ora_extractor_tables(void* h, float* scale, float* invScale, float* sigma, float* invSigma, int* patchSz,
                          int* quota, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nLevels; i++) {
        scale[i] = e->scalePyramid[i]; invScale[i] = e->scaleInvPyramid[i]; sigma[i] = e->sigmaFactor[i];
        invSigma[i] = e->InvSigmaFactor[i]; patchSz[i] = e->scaledPatchSize[i]; quota[i] = e->featurePerLevel[i];
    }
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

// returns number of keypoints; kps (cap x 28 B), desc (cap x 32 B); -1 if cap too small
int ora_extract(void* h, const uint8_t* img, int w, int hgt, int stride, void* kps, uint8_t* desc, int cap) {
    Extractor* e = (Extractor*)h;
    std::vector<KeyPoint> k; std::vector<uint8_t> d;
    int n = e->extract(img, w, hgt, stride, k, d);
    if (n > cap) return -1;
    if (n > 0) { std::memcpy(kps, k.data(), (size_t)n * sizeof(KeyPoint)); std::memcpy(desc, d.data(), (size_t)n * 32); }
    return n;
}

void ora_level_size(void* h, int level, int* w, int* hgt) {
    Extractor* e = (Extractor*)h; *w = e->imagePyramid[level].w; *hgt = e->imagePyramid[level].h;
}
void ora_level_image(void* h, int level, int blurred, uint8_t* out) {
    Extractor* e = (Extractor*)h;
    const Image& im = blurred ? e->blurred[level] : e->imagePyramid[level];
    if (!im.d.empty()) std::memcpy(out, im.d.data(), im.d.size());
}
// which: 0 = FAST candidates (pre-SSC, emission order), 1 = kept keys (post-SSC order)
int ora_level_count(void* h, int level, int which) {
    Extractor* e = (Extractor*)h; return (int)(which ? e->kept[level] : e->cands[level]).x.size();
}
void ora_level_keys(void* h, int level, int which, int* x, int* y, float* resp, float* angle, int* m10, int* m01) {
    Extractor* e = (Extractor*)h;
    const LevelKeys& k = which ? e->kept[level] : e->cands[level];
    size_t n = k.x.size();
    std::memcpy(x, k.x.data(), n * 4); std::memcpy(y, k.y.data(), n * 4); std::memcpy(resp, k.response.data(), n * 4);
    if (which && angle) { std::memcpy(angle, k.angle.data(), n * 4); std::memcpy(m10, k.m10.data(), n * 4); std::memcpy(m01, k.m01.data(), n * 4); }
}
int ora_level_ssc_passes(void* h, int level) { return ((Extractor*)h)->sscPasses[level]; }

void ora_stereo_match(void* hL, void* hR, double fx, float baseline, int imageHeight, const void* kl, const uint8_t* dl,
                      int NL, const void* kr, const uint8_t* dr, int NR, int* rightIdxs, int* leftIdxs, float* depth,
                      uint8_t* close, int* stats) {
    stereoMatch(*(Extractor*)hL, *(Extractor*)hR, fx, baseline, imageHeight, (const KeyPoint*)kl, dl, NL,
                (const KeyPoint*)kr, dr, NR, rightIdxs, leftIdxs, depth, close, stats);
}

// grid in CSR form: start[(yGrids*64)+1], items[n]; returns yGrids
int ora_assign_grid(const void* kps, int n, int width, int height, int* start, int* items, float* xMult, float* yMult) {
    Grid g = assignKeysToGrids((const KeyPoint*)kps, n, width, height);
    std::memcpy(start, g.start.data(), g.start.size() * 4);
    if (n) std::memcpy(items, g.items.data(), (size_t)n * 4);
    *xMult = g.xMult; *yMult = g.yMult;
    return g.yGrids;
}

// mode 0: matchByProjectionRPred (src/FeatureMatcher.cpp:254-389); 1: matchByProjectionMono (:391-456);
// 2: matchByRadius (:458-526; the "map points" are lastKeys' keypoints: predL = pt, scaleLevelL = octave,
//    desc = Desc row, matches[i].first receives the matched index instead of the KeyFrame* list).
// scalePyramid comes from feLeft.  leftIdxs/rightIdxs = the frame's stereo partner arrays (mode 0 only).
int ora_match_projection(void* hfe, int mode, int M, const uint8_t* mpDesc, const float* predL, const float* predR,
                         const int* lvlL, const int* lvlR, const uint8_t* inFrame, const uint8_t* inFrameR,
                         const void* kpL, const uint8_t* descL, int NL, const void* kpR, const uint8_t* descR, int NR,
                         const int* leftIdxs, const int* rightIdxs, int width, int height, float rad,
                         int* matchedIdxsL, int* matchedIdxsR, int* matchesIdxs /* M x 2 */) {
    Extractor* fe = (Extractor*)hfe;
    const KeyPoint* kl = (const KeyPoint*)kpL; const KeyPoint* kr = (const KeyPoint*)kpR;
    Grid gl = assignKeysToGrids(kl, NL, width, height);
    Grid gr;
    if (mode == 0) gr = assignKeysToGrids(kr, NR, width, height);
    const int matchDistProj = 100; const float ratioProj = 0.8f;
    int nMatches = 0;
    auto noskip = [](int) { return false; };
    for (int i = 0; i < M; i++) {
        int* keyPair = matchesIdxs + 2 * i;
        if (mode != 2 && (keyPair[0] >= 0 || keyPair[1] >= 0)) continue;
        const uint8_t* d = mpDesc + (size_t)i * 32;
        int lvl = lvlL[i];
        float radius = fe->scalePyramid[lvl] * rad;
        std::vector<int> idxs;
        getMatchIdxs(predL[2 * i], predL[2 * i + 1], idxs, gl, kl, lvl, radius);
        Best b;
        if (mode == 2) {
            const double kx = predL[2 * i], ky = predL[2 * i + 1];
            auto noParallax = [&](int idx) {  // include/Conversions.h:140-144
                double dx = (double)kl[idx].x - kx, dy = (double)kl[idx].y - ky;
                return !(std::sqrt(dx * dx + dy * dy) > 10.0);
            };
            if (!idxs.empty()) b = scanCandidates(d, idxs, matchedIdxsL, kl, descL, noParallax);
        } else if (!idxs.empty() && inFrame[i]) {
            b = scanCandidates(d, idxs, matchedIdxsL, kl, descL, noskip);
        }
        bool right = false;
        Best br;
        if (mode == 0) {
            std::vector<int> idxsR;
            lvl = lvlR[i];
            radius = fe->scalePyramid[lvl] * rad;
            getMatchIdxs(predR[2 * i], predR[2 * i + 1], idxsR, gr, kr, lvl, radius);
            if (!idxsR.empty() && inFrameR[i]) br = scanCandidates(d, idxsR, matchedIdxsR, kr, descR, noskip);
            if (b.bestDist > br.bestDist) {
                b.bestDist = br.bestDist; b.secDist = br.secDist; b.bestLev = br.bestLev; b.bestLev2 = br.bestLev2;
                right = true;
            }
        }
        if (mode == 1) {
            if (b.bestDist > (matchDistProj + 50)) continue;
            if (b.bestLev == b.bestLev2 && b.bestDist >= (ratioProj + 0.1) * b.secDist) continue;
            if (b.bestLev != b.bestLev2 || b.bestDist < (ratioProj + 0.1) * b.secDist) {
                nMatches++;
                matchedIdxsL[b.bestIdx] = i;
                keyPair[0] = b.bestIdx;
            }
            continue;
        }
        if (b.bestDist > matchDistProj) continue;
        if (b.bestLev == b.bestLev2 && b.bestDist >= ratioProj * b.secDist) continue;
        if (b.bestLev != b.bestLev2 || b.bestDist < ratioProj * b.secDist) {
            nMatches++;
            if (mode == 2) { matchedIdxsL[b.bestIdx] = i; keyPair[0] = b.bestIdx; continue; }
            if (right) {
                matchedIdxsR[br.bestIdx] = i;
                keyPair[1] = br.bestIdx;
                if (leftIdxs[br.bestIdx] >= 0) { keyPair[0] = leftIdxs[br.bestIdx]; matchedIdxsL[leftIdxs[br.bestIdx]] = i; }
            } else {
                matchedIdxsL[b.bestIdx] = i;
                keyPair[0] = b.bestIdx;
                if (rightIdxs[b.bestIdx] >= 0) { keyPair[1] = rightIdxs[b.bestIdx]; matchedIdxsR[rightIdxs[b.bestIdx]] = i; }
            }
        }
    }
    return nMatches;
}

// ssc alone on a caller-made candidate list (cols, rows = level size); returns kept count, kept indices in out
int ora_ssc(void* h, int numRetPoints, int cols, int rows, int n, const int* x, const int* y, const float* resp, int* out,
            int* sortedOut) {
    Extractor* e = (Extractor*)h;
    LevelKeys k;
    k.x.assign(x, x + n); k.y.assign(y, y + n); k.response.assign(resp, resp + n);
    if (sortedOut) sortIdxDescending(resp, n, sortedOut);
    std::vector<int> keep = e->ssc(k, numRetPoints, 0.1f, cols, rows, nullptr);
    std::memcpy(out, keep.data(), keep.size() * 4);
    return (int)keep.size();
}

// FeatureMatcher::matchByProjectionRPredLBA (src/FeatureMatcher.cpp:66-252) with the KeyFrame / Eigen inputs flattened
// (see oracle/ref_cv2.py: matchByProjectionRPredLBA for the mapping).  out: M x 2, (-2,-2) = nothing appended.
int ora_match_lba(void* hfe, int M, const uint8_t* mpDesc, const float* predL, const float* predR, const int* predScale,
                  const float* lastL, const float* lastR, const void* kpL, const uint8_t* descL, int NL, const void* kpR,
                  const uint8_t* descR, int NR, const int* leftIdxs, const int* rightIdxs, const int* unMatchedF,
                  const int* unMatchedFR, int width, int height, float rad, int* out) {
    Extractor* fe = (Extractor*)hfe;
    const KeyPoint* kl = (const KeyPoint*)kpL; const KeyPoint* kr = (const KeyPoint*)kpR;
    Grid gl = assignKeysToGrids(kl, NL, width, height), gr = assignKeysToGrids(kr, NR, width, height);
    const int matchDistLBA = 50; const float ratioLBA = 0.6f;
    int nMatches = 0;
    for (int i = 0; i < M; i++) {
        out[2 * i] = out[2 * i + 1] = -2;
        const uint8_t* d = mpDesc + (size_t)i * 32;
        const int ps = predScale[i];
        int bestDist = 256, bestIdx = -1, bestLev = -1, bestLev2 = -1, secDist = 256;
        float radius = fe->scalePyramid[ps] * rad;
        if (predL[2 * i] > 0 && predL[2 * i + 1] > 0) {
            std::vector<int> idxs;
            getMatchIdxs(predL[2 * i], predL[2 * i + 1], idxs, gl, kl, ps, radius);
            for (int idx : idxs) {
                if (unMatchedF[idx] >= 0) continue;
                const int lev = kl[idx].octave;
                int dist = descriptorDistance(d, descL + (size_t)idx * 32);
                if (dist < bestDist) { secDist = bestDist; bestLev2 = bestLev; bestDist = dist; bestLev = lev; bestIdx = idx; continue; }
                if (dist < secDist) { secDist = dist; bestLev2 = bestLev; }  // sic, :132-136
            }
        }
        radius = fe->scalePyramid[ps] * rad;
        int bestDistR = 256, bestIdxR = -1, bestLevR = -1, bestLevR2 = -1, secDistR = 256;
        if (predR[2 * i] > 0 && predR[2 * i + 1] > 0) {
            std::vector<int> idxs;
            getMatchIdxs(predR[2 * i], predR[2 * i + 1], idxs, gr, kr, ps, radius);
            for (int idx : idxs) {
                if (unMatchedFR[idx] >= 0) continue;
                const int lev = kr[idx].octave;
                int dist = descriptorDistance(d, descR + (size_t)idx * 32);
                if (dist < bestDistR) { secDistR = bestDistR; bestLevR2 = bestLevR; bestDistR = dist; bestLevR = lev; bestIdxR = idx; continue; }
                if (dist < secDistR) { secDistR = dist; bestLevR2 = lev; }
            }
        }
        bool right = false;
        if (bestDist > bestDistR) { bestDist = bestDistR; secDist = secDistR; bestLev = bestLevR; bestLev2 = bestLevR2; right = true; }
        if (bestDist > matchDistLBA) continue;
        if (bestLev == bestLev2 && bestDist >= ratioLBA * secDist) continue;
        if (bestLev != bestLev2 || bestDist < ratioLBA * secDist) {
            const float* last = right ? lastR + 2 * i : lastL + 2 * i;
            const float* pred = right ? predR + 2 * i : predL + 2 * i;
            const double dx = (double)pred[0] - (double)last[0], dy = (double)pred[1] - (double)last[1];
            if (!(std::sqrt(dx * dx + dy * dy) > 10.0)) continue;  // Converter::checkPixelParallax
            if (right) { out[2 * i] = leftIdxs[bestIdxR] >= 0 ? leftIdxs[bestIdxR] : -1; out[2 * i + 1] = bestIdxR; }
            else { out[2 * i] = bestIdx; out[2 * i + 1] = rightIdxs[bestIdx] >= 0 ? rightIdxs[bestIdx] : -1; }
            nMatches++;
        }
    }
    return nMatches;
}

// MapPoint::calcDescriptor (src/Map.cpp:145-210) for P map points: descs = concatenated observation descriptors,
// offsets[P+1]; bestIdx[p] = index (inside the point's own list) of the descriptor with the least median distance.
void ora_median_descriptor(const uint8_t* descs, const int* offsets, int P, int* bestIdx) {
    for (int p = 0; p < P; p++) {
        const uint8_t* d = descs + (size_t)offsets[p] * 32;
        const int N = offsets[p + 1] - offsets[p];
        int BestMedian = INT_MAX, BestIdx = 0;
        std::vector<int> v(N);
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) v[j] = i == j ? 0 : descriptorDistance(d + (size_t)i * 32, d + (size_t)j * 32);
            std::sort(v.begin(), v.end());
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        bestIdx[p] = N > 0 ? BestIdx : -1;
    }
}

// Upstream image preparation: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) with CV_32FC1 maps
// (src/VIOSlam.cpp:282-301) followed by cvtColor BGR2GRAY (src/FeatureTracker.cpp:1131-1135); channels = 1 or 3.
// OpenCV arithmetic (pinned against cv2 in tests): sx = cvRound(mapx * 32) (float32 product), 5 fractional bits,
// weights (32-fx)(32-fy)*32 ... saturated to short, value = (sum + 2^14) >> 15 per channel; gray = (3735 B + 19235 G +
// 9798 R + 2^14) >> 15.  mapx == NULL skips the remap.
void ora_prepare_image(const uint8_t* src, int w, int h, int channels, const float* mapx, const float* mapy, uint8_t* gray) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v[3] = {0, 0, 0};
            if (mapx) {
                const int sx = cvRoundF(mapx[(size_t)y * w + x] * 32.f), sy = cvRoundF(mapy[(size_t)y * w + x] * 32.f);
                const int fx = sx & 31, fy = sy & 31;
                const int ix = std::min(std::max(sx >> 5, -32768), 32767), iy = std::min(std::max(sy >> 5, -32768), 32767);
                const int wt[4] = {std::min((32 - fx) * (32 - fy) * 32, 32767), fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
                for (int c = 0; c < channels; c++) {
                    long acc = 0;
                    for (int t = 0; t < 4; t++) {
                        const int px = ix + (t & 1), py = iy + (t >> 1);
                        const int val = (px >= 0 && px < w && py >= 0 && py < h) ? src[((size_t)py * w + px) * channels + c] : 0;
                        acc += (long)val * wt[t];
                    }
                    v[c] = (int)std::min<long>(std::max<long>((acc + 16384) >> 15, 0), 255);
                }
            } else {
                for (int c = 0; c < channels; c++) v[c] = src[((size_t)y * w + x) * channels + c];
            }
            gray[(size_t)y * w + x] = channels == 3 ? (uint8_t)((v[0] * 3735 + v[1] * 19235 + v[2] * 9798 + 16384) >> 15) : (uint8_t)v[0];
        }
}

// FeatureTracker::worldToFrame (src/FeatureTracker.cpp:685-741) + MapPoint::predictScale (src/Map.cpp:13-23) for M map points
// and one camera.  poseInv: 16 doubles row-major; pred / scaleLevel in-out; inFrame out.  The 4x4 * vec4
// product is evaluated row by row, left to right, without contraction (the canonical definition; Eigen's own order with
// -march=native may fuse).
void ora_world_to_frame(const double* poseInv, double fx, double fy, double cx, double cy, int width, int height, int M,
                        const double* wPos, const float* maxScaleDist, const float* logScale, const int* nScaleLev, float* pred,
                        int* scaleLevel, uint8_t* inFrame) {
    for (int i = 0; i < M; i++) {
        const double* w = wPos + 4 * i;
        double pt[4];
        for (int r = 0; r < 4; r++) pt[r] = ((poseInv[4 * r] * w[0] + poseInv[4 * r + 1] * w[1]) + poseInv[4 * r + 2] * w[2]) + poseInv[4 * r + 3] * w[3];
        if (pt[2] <= 0.0) { inFrame[i] = 0; continue; }
        const double invZ = 1.0f / pt[2];
        const double u = fx * pt[0] * invZ + cx, v = fy * pt[1] * invZ + cy;
        if (u < 0 || v < 0 || u >= width || v >= height) { inFrame[i] = 0; continue; }
        const float dist = (float)std::sqrt(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]);
        const float dif = maxScaleDist[i] / dist;
        // unqualified `log(dif)` on a float (src/Map.cpp:16) binds to ::log(double) when only <cmath> is in scope
        int scale = (int)std::ceil(std::log((double)dif) / (double)logScale[i]);
        if (scale < 0) scale = 0; else if (scale >= nScaleLev[i]) scale = nScaleLev[i] - 1;
        scaleLevel[i] = scale; inFrame[i] = 1; pred[2 * i] = (float)u; pred[2 * i + 1] = (float)v;
    }
}

// FeatureTracker::check2dError (src/FeatureTracker.cpp:147-164), on the first three rows of pose * wp
static bool check2dErrorRef(const double* P, const double* w, float ox, float oy, double fx, double fy, double cx, double cy, double thres,
                            double weight, double* zOut) {
    double p[3];
    for (int r = 0; r < 3; r++) p[r] = ((P[4 * r] * w[0] + P[4 * r + 1] * w[1]) + P[4 * r + 2] * w[2]) + P[4 * r + 3] * w[3];
    if (zOut) *zOut = p[2];
    if (p[2] <= 0) return true;
    const double invZ = 1.0f / p[2];
    const double u = fx * p[0] * invZ + cx, v = fy * p[1] * invZ + cy;
    const double errorU = (double)ox - u, errorV = (double)oy - v;
    const double error = (errorU * errorU + errorV * errorV) * weight;
    return error > thres;
}

// FeatureTracker::findOutliersR (src/FeatureTracker.cpp:582-649; mode 0) / findOutliersMono (:651-684; mode 1), the sequential
// loop as written.  pose / poseR: estimPose / toCameraR, 16 doubles row-major (the two inverses of :584-585
// are the caller's).  matches (M x 2), outliers, depth, close, rightIdxs, leftIdxs are in-out; counts[0] += nInliers++,
// counts[1] = nStereo.  Restated, not compiled from the reference: FeatureTracker.cpp needs GTSAM.
void ora_find_outliers(int mode, const double* pose, const double* poseR, double fx, double fy, double cx, double cy, double thres,
                       float closeDist, const float* invSigma, int M, const uint8_t* valid, const double* wPos, const uint8_t* inFrame,
                       const uint8_t* inFrameR, int* matches, uint8_t* outliers, const KeyPoint* kl, const KeyPoint* kr, float* depth,
                       uint8_t* close, int* rightIdxs, int* leftIdxs, int* counts) {
    int nStereo = 0;
    for (int i = 0; i < M; i++) {
        if (valid && !valid[i]) continue;
        const double* w = wPos + 4 * i;
        int nIdx; bool right = false; const double* P;
        if (matches[2 * i] >= 0) {
            if (inFrame && !inFrame[i]) continue;
            P = pose; nIdx = matches[2 * i];
        } else if (mode == 0 && matches[2 * i + 1] >= 0) {
            if (inFrameR && !inFrameR[i]) continue;
            right = true; P = poseR; nIdx = matches[2 * i + 1];
        } else continue;
        const KeyPoint& k = right ? kr[nIdx] : kl[nIdx];
        double z = 0;
        const bool outlier = check2dErrorRef(P, w, k.x, k.y, fx, fy, cx, cy, thres, (double)invSigma[k.octave], &z);
        outliers[i] = outlier;
        if (outlier) continue;
        counts[0]++;
        if (mode != 0) continue;
        if (z < closeDist && close[nIdx] && !right) {
            if (matches[2 * i + 1] < 0) continue;
            const KeyPoint& k2 = kr[matches[2 * i + 1]];
            if (!check2dErrorRef(poseR, w, k2.x, k2.y, fx, fy, cx, cy, thres, (double)invSigma[k2.octave], nullptr)) nStereo++;
            else {
                depth[nIdx] = -1; close[nIdx] = 0;
                const int rIdx = rightIdxs[nIdx];
                rightIdxs[nIdx] = -1;
                if (rIdx >= 0) leftIdxs[rIdx] = -1;
                matches[2 * i + 1] = -1;
            }
        }
    }
    counts[1] = nStereo;
}

// FeatureTracker::removeOutOfFrameMPs (src/FeatureTracker.cpp:910-940; mode 0), removeOutOfFrameMPsMono (:942-967; mode 1) and
// PredictMPsPosition (:969-1013; mode 2) as sequential loops over ora_world_to_frame's single-point body.
// valid[i] = the reference does not `continue` for point i.  Restated, not compiled from the reference (FeatureTracker.cpp needs GTSAM).
void ora_predict_mappoints(int mode, const double* toCamera, const double* toCameraR, double fx, double fy, double cx, double cy, int width,
                           int height, int M, const uint8_t* valid, const double* wPos, const float* maxScaleDist, const float* logScale,
                           const int* nScaleLev, float* predL, int* scaleLevelL, uint8_t* inFrame, float* predR, int* scaleLevelR,
                           uint8_t* inFrameR, uint8_t* keep, int* matches, int* matchedL, int* matchedR, uint8_t* outliers) {
    for (int i = 0; i < M; i++) {
        if (mode != 2) keep[i] = 0;
        if (valid && !valid[i]) continue;
        uint8_t c1 = 0, c2 = 1;
        ora_world_to_frame(toCamera, fx, fy, cx, cy, width, height, 1, wPos + 4 * i, maxScaleDist + i, logScale + i, nScaleLev + i,
                           predL + 2 * i, scaleLevelL + i, &c1);
        inFrame[i] = c1;
        if (mode != 1) {
            ora_world_to_frame(toCameraR, fx, fy, cx, cy, width, height, 1, wPos + 4 * i, maxScaleDist + i, logScale + i, nScaleLev + i,
                               predR + 2 * i, scaleLevelR + i, &c2);
            inFrameR[i] = c2;
        }
        if (mode != 2) { keep[i] = c1 && c2; continue; }
        int& first = matches[2 * i]; int& second = matches[2 * i + 1];
        if (!c1 && first >= 0) { matchedL[first] = -1; first = -1; }
        if (!c2 && second >= 0) { matchedR[second] = -1; second = -1; }
        if (outliers[i]) {
            outliers[i] = 0;
            if (first >= 0) { matchedL[first] = -1; first = -1; }
            if (second >= 0) { matchedR[second] = -1; second = -1; }
        }
    }
}

// The map points the tracker holds for frame f when they all come from the stereo keys of one earlier frame (temporal
// matching of consecutive frames), as the reference builds and filters them:
//   FeatureTracker::initializeMap (src/FeatureTracker.cpp:89-112): one point per left key with estimatedDepth > 0,
//     p = pose * ((x - cx) * zp / fx, (y - cy) * zp / fy, zp, 1)  (double);
//   MapPoint::update (src/Map.cpp:61-104): dist = |wp3d - kF->pose.t| (float), maxScaleDist = dist * kF->scaleFactor[octave of the left key];
//   FeatureTracker::removeOutOfFrameMPs (:910-940): worldToFrame into both cameras, a point stays iff both succeed.
// Outputs, one entry per surviving point in order: mpKey (source key), predL / predR, scaleLevelL / R, wPos (4 doubles), maxScaleDist.
// Returns the number of surviving points.  Restated (FeatureTracker.cpp needs GTSAM); 4x4 products row by row, no contraction.
int ora_temporal_mappoints(const KeyPoint* kl, int NL, const float* depth, const double* posePrev, const double* toCamera,
                           const double* toCameraR, double fx, double fy, double cx, double cy, int width, int height,
                           const float* scalePyramid, int nLevels, float logScale, int* mpKey, float* predL, float* predR, int* lvlL,
                           int* lvlR, double* wPos, float* maxScaleDist) {
    int M = 0;
    for (int i = 0; i < NL; i++) {
        if (!(depth[i] > 0)) continue;
        const double zp = (double)depth[i];
        const double c[4] = {(double)(((double)kl[i].x - cx) * zp / fx), (double)(((double)kl[i].y - cy) * zp / fy), zp, 1.0};
        double w[4];
        for (int r = 0; r < 4; r++) w[r] = ((posePrev[4 * r] * c[0] + posePrev[4 * r + 1] * c[1]) + posePrev[4 * r + 2] * c[2]) + posePrev[4 * r + 3] * c[3];
        const double ex = w[0] - posePrev[3], ey = w[1] - posePrev[7], ez = w[2] - posePrev[11];
        const float dist = (float)std::sqrt(ex * ex + ey * ey + ez * ez);
        const float msd = dist * scalePyramid[kl[i].octave];
        const int nl = nLevels;
        float pl[2] = {0, 0}, pr[2] = {0, 0};
        int ll = 0, lr = 0;
        uint8_t c1 = 0, c2 = 0;
        ora_world_to_frame(toCamera, fx, fy, cx, cy, width, height, 1, w, &msd, &logScale, &nl, pl, &ll, &c1);
        ora_world_to_frame(toCameraR, fx, fy, cx, cy, width, height, 1, w, &msd, &logScale, &nl, pr, &lr, &c2);
        if (!(c1 && c2)) continue;
        mpKey[M] = i; predL[2 * M] = pl[0]; predL[2 * M + 1] = pl[1]; predR[2 * M] = pr[0]; predR[2 * M + 1] = pr[1];
        lvlL[M] = ll; lvlR[M] = lr;
        if (wPos) for (int r = 0; r < 4; r++) wPos[4 * M + r] = w[r];
        if (maxScaleDist) maxScaleDist[M] = msd;
        M++;
    }
    return M;
}

int ora_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptorDistance(a, b); }

// ---- primitive-level entry points (pinned one by one against cv2 in tests/test_oracle_vs_cv2.py)
void ora_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    resizeLinear(src, sw, sh, sw, dst, dw, dh, dw);
}
void ora_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst) { gaussianBlur7(src, w, h, w, dst, w); }
float ora_fast_atan2(float y, float x) { return fastAtan2(y, x); }
void ora_sort_idx_desc(const float* key, int n, int* idx) { sortIdxDescending(key, n, idx); }
int ora_fast_detect(const uint8_t* img, int w, int h, int threshold, int* xs, int* ys, int* scores, int cap) {
    std::vector<FastPt> out; std::vector<uint8_t> scratch;
    fastDetect(img, w, w, h, threshold, out, scratch);
    int n = std::min((int)out.size(), cap);
    for (int i = 0; i < n; i++) { xs[i] = out[i].x; ys[i] = out[i].y; scores[i] = out[i].score; }
    return (int)out.size();
}

// ---- CPU baseline: F stereo frames (extract L, extract R, stereo match, two grids = FeatureTracker.cpp:56-70), then, when `prev`
// is given, the temporal step of bench.py (map points of frame f from the stereo keys of frame prev[f], matchByProjectionRPred
// on fresh index arrays, src/FeatureTracker.cpp:1176-1205).
// shape 0 = "reference-shaped": frames in sequence, 2 extraction threads per frame + serial matcher (2 cores).
// shape 1 = frame-parallel over `threads` workers (each worker runs L, R, match serially).
// Returns total matched pairs (so the work cannot be optimised away); nKeysOut[f] = left keypoints of frame f.
long ora_track_frames(int nF, int nL, float s, int edge, int patch, int maxF, int minF, double fx, float baseline,
                      const uint8_t* L, const uint8_t* R, int frames, int w, int h, int shape, int threads, int* nKeysOut,
                      const int* prev, const double* posePrev, const double* toCamera, const double* toCameraR, double cfx, double cfy,
                      double ccx, double ccy, float rad, int* nTemporalOut) {
    std::atomic<long> total{0};
    struct FrameKeys { std::vector<KeyPoint> kl, kr; std::vector<uint8_t> dl, dr; std::vector<int> ridx, lidx; std::vector<float> dep; };
    std::vector<FrameKeys> all(prev ? frames : 0);
    auto one = [&](int f, bool twoThreads) {
        Extractor feL(nF, nL, s, edge, patch, maxF, minF), feR(nF, nL, s, edge, patch, maxF, minF);
        FrameKeys local;
        FrameKeys& k = prev ? all[f] : local;
        const uint8_t* li = L + (size_t)f * w * h; const uint8_t* ri = R + (size_t)f * w * h;
        int nl = 0, nr = 0;
        if (twoThreads) {
            std::thread a([&] { nl = feL.extract(li, w, h, w, k.kl, k.dl); });
            std::thread b([&] { nr = feR.extract(ri, w, h, w, k.kr, k.dr); });
            a.join(); b.join();
        } else { nl = feL.extract(li, w, h, w, k.kl, k.dl); nr = feR.extract(ri, w, h, w, k.kr, k.dr); }
        k.ridx.resize(nl); k.lidx.resize(nr); k.dep.resize(nl);
        std::vector<uint8_t> cl(nl);
        stereoMatch(feL, feR, fx, baseline, h, k.kl.data(), k.dl.data(), nl, k.kr.data(), k.dr.data(), nr, k.ridx.data(),
                    k.lidx.data(), k.dep.data(), cl.data(), nullptr);
        Grid g1 = assignKeysToGrids(k.kl.data(), nl, w, h), g2 = assignKeysToGrids(k.kr.data(), nr, w, h);
        long m = 0;
        for (int v : k.ridx) m += v >= 0;
        total += m + (g1.items.size() + g2.items.size() == (size_t)(nl + nr) ? 0 : 1);
        if (nKeysOut) nKeysOut[f] = nl;
    };
    auto temporal = [&](int f) {
        if (nTemporalOut) nTemporalOut[f] = 0;
        const int p = prev[f];
        if (p < 0) return;
        Extractor fe(nF, nL, s, edge, patch, maxF, minF);
        const FrameKeys& pk = all[p];
        const FrameKeys& k = all[f];
        const int np = (int)pk.kl.size(), cap = std::max(np, 1);
        std::vector<int> key(cap), ll(cap), lr(cap);
        std::vector<float> pl(2 * cap), pr(2 * cap);
        const int M = ora_temporal_mappoints(pk.kl.data(), np, pk.dep.data(), posePrev + 16 * f, toCamera + 16 * f, toCameraR + 16 * f, cfx, cfy,
                                             ccx, ccy, w, h, fe.scalePyramid.data(), nL, (float)std::log((double)s), key.data(), pl.data(),
                                             pr.data(), ll.data(), lr.data(), nullptr, nullptr);
        std::vector<uint8_t> desc((size_t)std::max(M, 1) * 32), in(std::max(M, 1), 1);
        for (int i = 0; i < M; i++) std::memcpy(&desc[(size_t)i * 32], &pk.dl[(size_t)key[i] * 32], 32);
        std::vector<int> mL(std::max<size_t>(k.kl.size(), 1), -1), mR(std::max<size_t>(k.kr.size(), 1), -1), mt(2 * (size_t)std::max(M, 1), -1);
        const int n = ora_match_projection(&fe, 0, M, desc.data(), pl.data(), pr.data(), ll.data(), lr.data(), in.data(), in.data(), k.kl.data(),
                                           k.dl.data(), (int)k.kl.size(), k.kr.data(), k.dr.data(), (int)k.kr.size(), k.lidx.data(), k.ridx.data(),
                                           w, h, rad, mL.data(), mR.data(), mt.data());
        total += n;
        if (nTemporalOut) nTemporalOut[f] = n;
    };
    auto parallel = [&](auto&& fn) {
        std::atomic<int> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < std::max(1, threads); t++)
            pool.emplace_back([&] { for (int f; (f = next++) < frames;) fn(f); });
        for (auto& t : pool) t.join();
    };
    if (shape == 0) {
        for (int f = 0; f < frames; f++) one(f, true);
        if (prev) for (int f = 0; f < frames; f++) temporal(f);
    } else {
        parallel([&](int f) { one(f, false); });
        if (prev) parallel(temporal);
    }
    return total.load();
}

long ora_stereo_frames(int nF, int nL, float s, int edge, int patch, int maxF, int minF, double fx, float baseline,
                       const uint8_t* L, const uint8_t* R, int frames, int w, int h, int shape, int threads,
                       int* nKeysOut) {
    return ora_track_frames(nF, nL, s, edge, patch, maxF, minF, fx, baseline, L, R, frames, w, h, shape, threads, nKeysOut, nullptr, nullptr,
                            nullptr, nullptr, 0, 0, 0, 0, 0.f, nullptr);
}

}  // extern "C"
