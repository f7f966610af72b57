"""TEST INFRASTRUCTURE ONLY -- CPU oracle #1: the reference's frontend hot path restated on the
real OpenCV primitives (Python cv2), function by function.

Nothing under gtsam_vslam_b200/ may import this file; only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py do.

Parity status: PINNED.  The reference (/root/reference) has no tests, golden vectors or fixtures of its own
(SURVEY.md section 4), but its two source files on the path compile here against stand-in headers
(oracle/_ref, see oracle/ref_build.py) and reproduce this file's fixtures bit for bit when their OpenCV
primitives are hooked to cv2 (tests/test_reference_build.py).  In this file every arithmetic
primitive on the path (cv::resize, cv::FAST, cv::sortIdx, cv::fastAtan2, cv::GaussianBlur, cv::norm)
is executed by the same OpenCV code the reference links (cv2 4.13.0 here; the reference asks for
>= 4.2), and the control flow around them follows the cited reference lines.  oracle/cpu_ref.cpp
(scalar C++, no OpenCV) is validated against this file and against tests/golden/*.npz made by it.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import cv2
import numpy as np

cv2.setNumThreads(1)

_F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_pattern() -> np.ndarray:
    txt = open(os.path.join(_HERE, "..", "include", "gvs_orb_pattern.inc")).read()
    txt = txt[txt.index("*/") + 2:]
    v = np.array([int(t) for t in txt.replace("\n", "").split(",") if t.strip()], dtype=np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)  # 512 points (x, y); test i uses points 2i, 2i+1


PATTERN = _load_pattern()


def cv_round(x) -> int:
    """cvRound: round half to even (SSE cvtsd2si); Python's round() is half-even too."""
    return int(round(float(x)))


def c_round(x) -> float:
    """std::round: half away from zero."""
    x = float(x)
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


# ----------------------------------------------------------------------------------------------
# FeatureExtractor  (include/FeatureExtractor.h:52-97, src/FeatureExtractor.cpp:620-682)
# ----------------------------------------------------------------------------------------------
class FeatureExtractor:
    def __init__(self, nFeatures=2000, nLevels=8, imScale=1.2, edgeThreshold=19, patchSize=31,
                 maxFastThreshold=20, minFastThreshold=7):
        # src/FeatureExtractor.cpp:620-644 -- float32 cumulative tables
        self.nFeatures, self.nLevels = int(nFeatures), int(nLevels)
        self.imScale = _F32(imScale)
        self.edgeThreshold, self.patchSize = int(edgeThreshold), int(patchSize)
        self.halfPatchSize = 15
        self.maxFastThreshold, self.minFastThreshold = int(maxFastThreshold), int(minFastThreshold)
        L = self.nLevels
        self.scalePyramid = np.ones(L, _F32)
        self.scaledPatchSize = np.full(L, self.patchSize, np.int32)
        self.sigmaFactor = np.ones(L, _F32)
        for i in range(1, L):
            self.scalePyramid[i] = _F32(self.scalePyramid[i - 1] * self.imScale)
            self.scaledPatchSize[i] = int(_F32(_F32(self.patchSize) * self.scalePyramid[i]))
            self.sigmaFactor[i] = _F32(self.scalePyramid[i] * self.scalePyramid[i])
        self.scaleInvPyramid = (_F32(1.0) / self.scalePyramid).astype(_F32)
        self.InvSigmaFactor = (_F32(1.0) / self.sigmaFactor).astype(_F32)
        # :648-659 per-level quota
        factor = _F32(_F32(1.0) / self.imScale)
        denom = _F32(_F32(1.0) - _F32(math.pow(float(factor), float(L))))
        nd = _F32(_F32(_F32(self.nFeatures) * _F32(_F32(1.0) - factor)) / denom)
        self.featurePerLevel = np.zeros(L, np.int32)
        s = 0
        for lvl in range(L - 1):
            self.featurePerLevel[lvl] = cv_round(nd)
            s += int(self.featurePerLevel[lvl])
            nd = _F32(nd * factor)
        self.featurePerLevel[L - 1] = max(self.nFeatures - s, 0)
        # :666-680 umax
        hp = self.halfPatchSize
        umax = [0] * (hp + 1)
        vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1))
        vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp * hp - v * v))
        v0 = 0
        for v in range(hp, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.imagePyramid: list[np.ndarray] = []

    # src/FeatureExtractor.cpp:342-366 (the 19-px copyMakeBorder frame is never read: SURVEY App. B.1)
    def computePyramid(self, image: np.ndarray) -> None:
        h, w = image.shape
        pyr = []
        for level in range(self.nLevels):
            scale = self.scaleInvPyramid[level]
            sz = (cv_round(_F32(_F32(w) * scale)), cv_round(_F32(_F32(h) * scale)))
            if level == 0:
                pyr.append(np.ascontiguousarray(image))
            else:
                pyr.append(cv2.resize(pyr[level - 1], sz, interpolation=cv2.INTER_LINEAR))
        self.imagePyramid = pyr

    # src/FeatureExtractor.cpp:315-340
    def computeOrientation(self, image: np.ndarray, x: float, y: float) -> float:
        cx, cy = cv_round(x), cv_round(y)
        hp = self.halfPatchSize
        img = image.astype(np.int32)
        m_10 = int(sum(u * img[cy, cx + u] for u in range(-hp, hp + 1)))
        m_01 = 0
        for v in range(1, hp + 1):
            d = self.umax[v]
            plus = img[cy + v, cx - d:cx + d + 1]
            minus = img[cy - v, cx - d:cx + d + 1]
            us = np.arange(-d, d + 1)
            m_01 += v * int((plus - minus).sum())
            m_10 += int((us * (plus + minus)).sum())
        return float(cv2.fastAtan2(float(m_01), float(m_10))), m_10, m_01

    # src/FeatureExtractor.cpp:368-468
    def ssc(self, pts: np.ndarray, resp: np.ndarray, numRetPoints: int, tolerance, cols: int, rows: int,
            stats: dict | None = None):
        """pts: (n,2) int (x,y); resp: (n,) float32.  Returns indices into the *input* order, in output order."""
        n = len(resp)
        idx = cv2.sortIdx(resp.reshape(1, -1).astype(np.float32), cv2.SORT_DESCENDING + cv2.SORT_EVERY_ROW)[0]
        sx = pts[idx, 0].astype(np.float64)
        sy = pts[idx, 1].astype(np.float64)
        exp1 = rows + cols + 2 * numRetPoints
        exp2 = (4 * cols + 4 * numRetPoints + 4 * rows * numRetPoints + rows * rows + cols * cols
                - 2 * rows * cols + 4 * rows * cols * numRetPoints)
        exp3 = math.sqrt(float(exp2))
        exp4 = float(numRetPoints - 1)
        sol1 = -c_round((exp1 + exp3) / exp4)
        sol2 = -c_round((exp1 - exp3) / exp4)
        high = int(sol1 if sol1 > sol2 else sol2)
        low = int(math.floor(math.sqrt(float(n) / numRetPoints)))
        low = max(1, low)
        prevWidth = -1
        K = numRetPoints
        tol = _F32(tolerance)
        Kmin = int(c_round(_F32(_F32(K) - _F32(_F32(K) * tol))))
        Kmax = int(c_round(_F32(_F32(K) + _F32(_F32(K) * tol))))
        result: list[int] = []
        ResultVec: list[int] = []
        passes = 0
        while True:
            width = low + int((high - low) / 2)  # C++ int division truncates toward zero
            if width == prevWidth or low > high:
                ResultVec = result
                break
            result = []
            c = width / 2.0
            numCellCols = int(math.floor(cols / c))
            numCellRows = int(math.floor(rows / c))
            covered = np.zeros((numCellRows + 1, numCellCols + 1), bool)
            rr = np.floor(sy / c).astype(np.int64)
            cc = np.floor(sx / c).astype(np.int64)
            reach = int(math.floor(width / c))
            passes += 1
            for i in range(n):
                row, col = rr[i], cc[i]
                if not covered[row, col]:
                    result.append(i)
                    covered[max(row - reach, 0):min(row + reach, numCellRows) + 1,
                            max(col - reach, 0):min(col + reach, numCellCols) + 1] = True
            if Kmin <= len(result) <= Kmax:
                ResultVec = result
                break
            elif len(result) < Kmin:
                high = width - 1
            else:
                low = width + 1
            prevWidth = width
        if stats is not None:
            stats.setdefault("ssc_passes", []).append(passes)
        return idx[np.array(ResultVec, dtype=np.int64)] if ResultVec else np.zeros(0, np.int64)

    # src/FeatureExtractor.cpp:535-618.  Returns per level: dict(x, y (int), response, angle)
    def computeKeypointsORBNew(self, stats: dict | None = None):
        fastEdge = 3
        W = _F32(35)
        minX = minY = self.edgeThreshold - fastEdge
        fast_hi = cv2.FastFeatureDetector_create(self.maxFastThreshold, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        fast_lo = cv2.FastFeatureDetector_create(self.minFastThreshold, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        allKeys = []
        for level in range(self.nLevels):
            img = self.imagePyramid[level]
            rows, cols = img.shape
            maxX = cols - self.edgeThreshold + fastEdge
            maxY = rows - self.edgeThreshold + fastEdge
            wid, hig = maxX - minX, maxY - minY
            nCols = int(_F32(wid) / W)
            nRows = int(_F32(hig) / W)
            gridW = int(math.ceil(_F32(_F32(wid) / _F32(nCols))))
            gridH = int(math.ceil(_F32(_F32(hig) / _F32(nRows))))
            xs, ys, rs = [], [], []
            for iR in range(nRows):
                rStart = minY + iR * gridH
                rEnd = rStart + gridH + 2 * fastEdge
                if rStart >= maxY - 2 * fastEdge:
                    continue
                rEnd = min(rEnd, maxY)
                for iC in range(nCols):
                    cStart = minX + iC * gridW
                    cEnd = cStart + gridW + 2 * fastEdge
                    if cStart >= maxX - 2 * fastEdge:
                        continue
                    cEnd = min(cEnd, maxX)
                    sub = img[rStart:rEnd, cStart:cEnd]
                    kps = fast_hi.detect(sub)
                    if len(kps) == 0:
                        kps = fast_lo.detect(sub)
                    for kp in kps:
                        xs.append(int(kp.pt[0]) + cStart)
                        ys.append(int(kp.pt[1]) + rStart)
                        rs.append(kp.response)
            pts = np.stack([np.array(xs, np.int32), np.array(ys, np.int32)], axis=1) if xs else np.zeros((0, 2), np.int32)
            resp = np.array(rs, np.float32)
            if stats is not None:
                stats.setdefault("candidates", []).append(len(resp))
            fPLevel = int(self.featurePerLevel[level])
            if len(resp) > fPLevel:
                keep = self.ssc(pts, resp, fPLevel, 0.1, cols, rows, stats)
                pts, resp = pts[keep], resp[keep]
            ang = np.zeros(len(resp), np.float32)
            for i in range(len(resp)):
                ang[i] = self.computeOrientation(img, pts[i, 0], pts[i, 1])[0]
            allKeys.append(dict(pts=pts, response=resp, angle=ang))
        return allKeys

    # src/FeatureExtractor.cpp:268-313
    def computeDescriptors(self, blurred: np.ndarray, pts: np.ndarray, angle: np.ndarray) -> np.ndarray:
        factorPI = _F32(math.pi / 180.0)
        n = len(angle)
        desc = np.zeros((n, 32), np.uint8)
        px = PATTERN[:, 0].astype(_F32)
        py = PATTERN[:, 1].astype(_F32)
        for i in range(n):
            ang = _F32(_F32(angle[i]) * factorPI)
            a = _F32(math.cos(float(ang)))
            b = _F32(math.sin(float(ang)))
            cx, cy = int(pts[i, 0]), int(pts[i, 1])
            # cvRound(x*b + y*a), cvRound(x*a - y*b) in float32 without contraction
            ry = np.rint((px * b).astype(_F32) + (py * a).astype(_F32)).astype(np.int32)
            rx = np.rint((px * a).astype(_F32) - (py * b).astype(_F32)).astype(np.int32)
            v = blurred[cy + ry, cx + rx].astype(np.int32)
            bits = (v[0::2] < v[1::2]).astype(np.uint8).reshape(32, 8)
            desc[i] = (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)
        return desc

    # src/FeatureExtractor.cpp:481-533.  Returns keypoints as a structured (N,7) set of arrays + desc (N,32)
    def extractKeysNew(self, image: np.ndarray, stats: dict | None = None):
        self.computePyramid(image)
        allKeys = self.computeKeypointsORBNew(stats)
        nF = sum(len(k["response"]) for k in allKeys)
        kp = np.zeros(nF, dtype=KEYPOINT_DTYPE)
        desc = np.zeros((nF, 32), np.uint8)
        if nF <= 0:
            return kp, desc
        o = 0
        for level in range(self.nLevels):
            k = allKeys[level]
            n = len(k["response"])
            if n <= 0:
                continue
            scale = self.scalePyramid[level]
            im = cv2.GaussianBlur(self.imagePyramid[level].copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
            d = self.computeDescriptors(im, k["pts"], k["angle"])
            x = k["pts"][:, 0].astype(_F32)
            y = k["pts"][:, 1].astype(_F32)
            if level != 0:
                x = (x * scale).astype(_F32)
                y = (y * scale).astype(_F32)
            kp["x"][o:o + n] = x
            kp["y"][o:o + n] = y
            kp["size"][o:o + n] = _F32(self.scaledPatchSize[level])
            kp["angle"][o:o + n] = k["angle"]
            kp["response"][o:o + n] = k["response"]
            kp["octave"][o:o + n] = level
            kp["class_id"][o:o + n] = -1
            desc[o:o + n] = d
            o += n
        return kp, desc


# cv::KeyPoint memory layout (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


# ----------------------------------------------------------------------------------------------
# TrackedKeys (include/FeatureExtractor.h:18-50)
# ----------------------------------------------------------------------------------------------
@dataclass
class TrackedKeys:
    keyPoints: np.ndarray = None
    rightKeyPoints: np.ndarray = None
    Desc: np.ndarray = None
    rightDesc: np.ndarray = None
    lkeyGrid: list = field(default_factory=list)
    rkeyGrid: list = field(default_factory=list)
    rightIdxs: np.ndarray = None
    leftIdxs: np.ndarray = None
    estimatedDepth: np.ndarray = None
    close: np.ndarray = None
    xMult: float = 0.0
    yMult: float = 0.0
    xGrids: int = 0
    yGrids: int = 0


def DescriptorDistance(a: np.ndarray, b: np.ndarray) -> int:
    """src/FeatureMatcher.cpp:710-726 (256-bit Hamming)."""
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


_POP8 = np.array([bin(i).count("1") for i in range(256)], np.int32)


def _hamming_many(a: np.ndarray, B: np.ndarray) -> np.ndarray:
    return _POP8[np.bitwise_xor(B, a[None, :])].sum(axis=1)


# ----------------------------------------------------------------------------------------------
# FeatureMatcher (include/FeatureMatcher.h:22-65)
# ----------------------------------------------------------------------------------------------
class FeatureMatcher:
    thDist = 75
    matchDistProj = 100
    ratioProj = _F32(0.8)
    closeNumber = 40

    def __init__(self, fx: float, baseline: float, feLeft: FeatureExtractor, feRight: FeatureExtractor,
                 imageHeight: int = 360):
        self.fx = float(fx)              # Camera::fx is double (include/Camera.h:72)
        self.mBaseline = _F32(baseline)  # StereoCamera::mBaseline is float (:86)
        self.feLeft, self.feRight, self.imageHeight = feLeft, feRight, int(imageHeight)

    # src/FeatureMatcher.cpp:728-752
    def destributeRightKeys(self, rightKeys: np.ndarray):
        indexes = [[] for _ in range(self.imageHeight)]
        for count in range(len(rightKeys)):
            yKey = cv_round(rightKeys["y"][count])
            r = _F32(_F32(2.0) * self.feRight.scalePyramid[rightKeys["octave"][count]])
            mn = int(math.floor(_F32(_F32(yKey) - r)))
            mx = int(math.ceil(_F32(_F32(yKey) + r)))
            for pos in range(mn, mx + 1):
                indexes[pos].append(count)
        return indexes

    # src/FeatureMatcher.cpp:528-708
    def findStereoMatchesORB2R(self, rightDesc: np.ndarray, rightKeys: np.ndarray, keysLeft: TrackedKeys,
                               stats: dict | None = None) -> None:
        indexes = self.destributeRightKeys(rightKeys)
        NL, NR = len(keysLeft.keyPoints), len(rightKeys)
        keysLeft.estimatedDepth = np.full(NL, -1.0, _F32)
        keysLeft.close = np.zeros(NL, bool)
        keysLeft.rightIdxs = np.full(NL, -1, np.int32)
        keysLeft.leftIdxs = np.full(NR, -1, np.int32)
        minD = _F32(0)
        maxD = _F32(self.fx)
        windowRadius, windowMovementX = 5, 5
        allDepths, allDists2 = [], []
        kl = keysLeft.keyPoints
        nham = 0
        for leftRow in range(NL):
            ptx, pty = _F32(kl["x"][leftRow]), _F32(kl["y"][leftRow])
            yKey = cv_round(pty)
            uL = pty                      # sic: pt.y, :557
            octL = int(kl["octave"][leftRow])
            minU = _F32(uL - maxD)
            maxU = _F32(uL - minD)
            if maxU < 0:
                continue
            bestDist, bestIdx = 256, -1
            cand = indexes[yKey]
            if len(cand) == 0:
                continue
            for idx in cand:
                uR = _F32(rightKeys["y"][idx])
                octR = int(rightKeys["octave"][idx])
                if octR < octL - 1 or octR > octL + 1:
                    continue
                if not (uR >= minU and uR <= maxU):
                    continue
                dist = DescriptorDistance(keysLeft.Desc[leftRow], rightDesc[idx])
                nham += 1
                if bestDist > dist:
                    bestDist, bestIdx = dist, idx
            if bestDist > self.thDist:
                continue
            kRx = _F32(rightKeys["x"][bestIdx])
            scale = self.feLeft.scaleInvPyramid[octL]
            scuL = c_round(_F32(ptx * scale))
            scvL = c_round(_F32(pty * scale))
            scuR = c_round(_F32(kRx * scale))
            imL = self.feLeft.imagePyramid[octL]
            imR = self.feRight.imagePyramid[octL]
            r0, r1 = int(scvL - windowRadius), int(scvL + windowRadius + 1)
            lWin = imL[r0:r1, int(scuL - windowRadius):int(scuL + windowRadius + 1)]
            bestDistW = 2147483647
            bestX = 0
            allDists = [_F32(0.0)] * (2 * windowMovementX + 1)
            for xMov in range(-windowMovementX, windowMovementX + 1):
                startW = scuR + xMov - windowRadius
                endW = scuR + xMov + windowRadius + 1
                if startW < 0 or endW >= imR.shape[1]:
                    continue
                rWin = imR[r0:r1, int(startW):int(endW)]
                dist = _F32(cv2.norm(lWin, rWin, cv2.NORM_L1))
                if _F32(bestDistW) > dist:
                    bestX = xMov
                    bestDistW = int(dist)
                allDists[xMov + windowMovementX] = dist
            if bestX == -windowMovementX or bestX == windowMovementX:
                continue
            dist1 = allDists[windowMovementX + bestX - 1]
            dist2 = allDists[windowMovementX + bestX]
            dist3 = allDists[windowMovementX + bestX + 1]
            with np.errstate(all="ignore"):
                den = _F32(_F32(2.0) * _F32(_F32(dist1 + dist3) - _F32(_F32(2.0) * dist2)))
                delta = _F32(_F32(dist1 - dist3) / den)
            if delta > 1 or delta < -1:
                continue
            newuR = _F32(self.feLeft.scalePyramid[octL] * _F32(_F32(_F32(scuR) + _F32(bestX)) + delta))
            disparity = _F32(ptx - newuR)
            if disparity > _F32(0.0) and float(disparity) < self.fx:
                depth = _F32(_F32(_F32(self.fx) * self.mBaseline) / disparity)
                keysLeft.rightIdxs[leftRow] = bestIdx
                keysLeft.leftIdxs[bestIdx] = leftRow
                keysLeft.estimatedDepth[leftRow] = depth
                allDists2.append((bestDistW, leftRow))
                allDepths.append((float(depth), leftRow))
                if depth < _F32(self.mBaseline * _F32(self.closeNumber)):
                    keysLeft.close[leftRow] = True
        if stats is not None:
            stats["hamming_tests"] = nham
            stats["accepted"] = len(allDepths)
        if len(allDepths) <= 0:
            return
        allDepths.sort()
        allDists2.sort()
        medianD = allDists2[len(allDists2) // 2][0]
        medDistD = _F32(_F32(medianD) * _F32(_F32(1.5) * _F32(1.4)))
        endDe = int(math.floor(len(allDepths) * 0.01))
        for i in range(endDe):
            li = allDepths[i][1]
            rIdx = keysLeft.rightIdxs[li]
            if rIdx >= 0:
                keysLeft.leftIdxs[rIdx] = -1
            keysLeft.rightIdxs[li] = -1
            keysLeft.estimatedDepth[li] = -1
            keysLeft.close[li] = False
        for i in range(len(allDists2) - 1, -1, -1):
            if _F32(allDists2[i][0]) < medDistD:
                break
            li = allDists2[i][1]
            rIdx = keysLeft.rightIdxs[li]
            if rIdx >= 0:
                keysLeft.leftIdxs[rIdx] = -1
            keysLeft.rightIdxs[li] = -1
            keysLeft.estimatedDepth[li] = -1
            keysLeft.close[li] = False

    # src/FeatureMatcher.cpp:13-64
    def getMatchIdxs(self, predP, keysLeft: TrackedKeys, predictedScale: int, radius, right: bool):
        idxs = []
        trackX, trackY = _F32(predP[0]), _F32(predP[1])
        radius = _F32(radius)
        xM, yM = _F32(keysLeft.xMult), _F32(keysLeft.yMult)
        with np.errstate(all="ignore"):
            minX = max(0, int(math.floor(_F32(_F32(trackX - radius) * xM))))
            maxX = min(keysLeft.xGrids - 1, int(math.ceil(_F32(_F32(trackX + radius) * xM))))
            minY = max(0, int(math.floor(_F32(_F32(trackY - radius) * yM))))
            maxY = min(keysLeft.yGrids - 1, int(math.ceil(_F32(_F32(trackY + radius) * yM))))
        if minX >= keysLeft.xGrids or minY >= keysLeft.yGrids or maxX < 0 or maxY < 0:
            return idxs
        maxLevel, minLevel = predictedScale + 1, predictedScale - 1
        grid = keysLeft.rkeyGrid if right else keysLeft.lkeyGrid
        kps = keysLeft.rightKeyPoints if right else keysLeft.keyPoints
        for row in range(minY, maxY + 1):
            for col in range(minX, maxX + 1):
                for gi in grid[row][col]:
                    o = kps["octave"][gi]
                    if o > maxLevel or o < minLevel:
                        continue
                    distx = _F32(kps["x"][gi] - trackX)
                    disty = _F32(kps["y"][gi] - trackY)
                    if abs(distx) < radius and abs(disty) < radius:
                        idxs.append(int(gi))
        return idxs

    def _scan(self, mpDesc, idxs, matched, kps, descs, extra_skip=None):
        """Best / second-best tracking shared by :281-302, :321-342, :418-439, :483-510."""
        bestDist, bestIdx, bestLev, bestLev2, secDist = 256, -1, -1, -1, 256
        for idx in idxs:
            if matched[idx] >= 0:
                continue
            if extra_skip is not None and extra_skip(idx):
                continue
            lev = int(kps["octave"][idx])
            dist = DescriptorDistance(mpDesc, descs[idx])
            if dist < bestDist:
                secDist, bestLev2 = bestDist, bestLev
                bestDist, bestLev, bestIdx = dist, lev, idx
                continue
            if dist < secDist:
                secDist, bestLev2 = dist, lev
        return bestDist, bestIdx, bestLev, bestLev2, secDist

    # src/FeatureMatcher.cpp:254-389.  mps: dict of arrays desc(M,32) predL(M,2) predR(M,2) scaleLevelL/R inFrame inFrameR
    def matchByProjectionRPred(self, mps: dict, keysLeft: TrackedKeys, matchedIdxsL, matchedIdxsR, matchesIdxs, rad):
        nMatches = 0
        rad = _F32(rad)
        for i in range(len(mps["desc"])):
            if matchesIdxs[i, 0] >= 0 or matchesIdxs[i, 1] >= 0:
                continue
            mpDesc = mps["desc"][i]
            lvl = int(mps["scaleLevelL"][i])
            radius = _F32(self.feLeft.scalePyramid[lvl] * rad)
            idxs = self.getMatchIdxs(mps["predL"][i], keysLeft, lvl, radius, False)
            bestDist, bestIdx, bestLev, bestLev2, secDist = 256, -1, -1, -1, 256
            if idxs and mps["inFrame"][i]:
                bestDist, bestIdx, bestLev, bestLev2, secDist = self._scan(
                    mpDesc, idxs, matchedIdxsL, keysLeft.keyPoints, keysLeft.Desc)
            lvl = int(mps["scaleLevelR"][i])
            radius = _F32(self.feLeft.scalePyramid[lvl] * rad)
            idxsR = self.getMatchIdxs(mps["predR"][i], keysLeft, lvl, radius, True)
            bestDistR, bestIdxR, bestLevR, bestLevR2, secDistR = 256, -1, -1, -1, 256
            if idxsR and mps["inFrameR"][i]:
                bestDistR, bestIdxR, bestLevR, bestLevR2, secDistR = self._scan(
                    mpDesc, idxsR, matchedIdxsR, keysLeft.rightKeyPoints, keysLeft.rightDesc)
            right = False
            if bestDist > bestDistR:
                bestDist, secDist, bestLev, bestLev2, right = bestDistR, secDistR, bestLevR, bestLevR2, True
            if bestDist > self.matchDistProj:
                continue
            thr = _F32(self.ratioProj * _F32(secDist))     # float * int -> float
            if bestLev == bestLev2 and _F32(bestDist) >= thr:
                continue
            if bestLev != bestLev2 or _F32(bestDist) < thr:
                nMatches += 1
                if right:
                    matchedIdxsR[bestIdxR] = i
                    matchesIdxs[i, 1] = bestIdxR
                    if keysLeft.leftIdxs[bestIdxR] >= 0:
                        matchesIdxs[i, 0] = keysLeft.leftIdxs[bestIdxR]
                        matchedIdxsL[keysLeft.leftIdxs[bestIdxR]] = i
                else:
                    matchedIdxsL[bestIdx] = i
                    matchesIdxs[i, 0] = bestIdx
                    if keysLeft.rightIdxs[bestIdx] >= 0:
                        matchesIdxs[i, 1] = keysLeft.rightIdxs[bestIdx]
                        matchedIdxsR[keysLeft.rightIdxs[bestIdx]] = i
        return nMatches

    # src/FeatureMatcher.cpp:391-456
    def matchByProjectionMono(self, mps: dict, keysLeft: TrackedKeys, matchedIdxsL, matchesIdxs, rad):
        nMatches = 0
        rad = _F32(rad)
        for i in range(len(mps["desc"])):
            if matchesIdxs[i, 0] >= 0 or matchesIdxs[i, 1] >= 0:
                continue
            lvl = int(mps["scaleLevelL"][i])
            radius = _F32(self.feLeft.scalePyramid[lvl] * rad)
            idxs = self.getMatchIdxs(mps["predL"][i], keysLeft, lvl, radius, False)
            bestDist, bestIdx, bestLev, bestLev2, secDist = 256, -1, -1, -1, 256
            if idxs and mps["inFrame"][i]:
                bestDist, bestIdx, bestLev, bestLev2, secDist = self._scan(
                    mps["desc"][i], idxs, matchedIdxsL, keysLeft.keyPoints, keysLeft.Desc)
            if bestDist > self.matchDistProj + 50:
                continue
            thr = (float(self.ratioProj) + 0.1) * float(secDist)   # float + double -> double (:445)
            if bestLev == bestLev2 and float(bestDist) >= thr:
                continue
            if bestLev != bestLev2 or float(bestDist) < thr:
                nMatches += 1
                matchedIdxsL[bestIdx] = i
                matchesIdxs[i, 0] = bestIdx
        return nMatches

    # src/FeatureMatcher.cpp:458-526.  Returns (nMatches, match_of_last[i] = matched index in actKeys or -1)
    def matchByRadius(self, lastKeys: TrackedKeys, actKeys: TrackedKeys, matchedIdxsL, rad):
        nMatches = 0
        rad = _F32(rad)
        out = np.full(len(lastKeys.keyPoints), -1, np.int32)
        for i in range(len(lastKeys.keyPoints)):
            key = lastKeys.keyPoints[i]
            lvl = int(key["octave"])
            radius = _F32(self.feLeft.scalePyramid[lvl] * rad)
            idxs = self.getMatchIdxs((key["x"], key["y"]), actKeys, lvl, radius, False)
            kx, ky = float(key["x"]), float(key["y"])

            def no_parallax(idx, kx=kx, ky=ky):
                # Converter::checkPixelParallax, include/Conversions.h:140-144: (a-b).norm() > 10 (double)
                dx = float(actKeys.keyPoints["x"][idx]) - kx
                dy = float(actKeys.keyPoints["y"][idx]) - ky
                return not (math.sqrt(dx * dx + dy * dy) > 10.0)

            bestDist, bestIdx, bestLev, bestLev2, secDist = 256, -1, -1, -1, 256
            if idxs:
                bestDist, bestIdx, bestLev, bestLev2, secDist = self._scan(
                    lastKeys.Desc[i], idxs, matchedIdxsL, actKeys.keyPoints, actKeys.Desc, no_parallax)
            if bestDist > self.matchDistProj:
                continue
            thr = _F32(self.ratioProj * _F32(secDist))
            if bestLev == bestLev2 and _F32(bestDist) >= thr:
                continue
            if bestLev != bestLev2 or _F32(bestDist) < thr:
                nMatches += 1
                matchedIdxsL[bestIdx] = i
                out[i] = bestIdx
        return nMatches, out


    # src/FeatureMatcher.cpp:66-252 (LocalMapper's matcher).  The KeyFrame / MapPoint / Eigen inputs are flattened:
    #   pts: desc (M,32) = mp->desc or the last keyframe's descriptor row (:76-86); predL/predR (M,2) = predPoints;
    #        predScale (M,) = the clamped cvCeil(log(dif)/logScale) of :88-95; lastL/lastR (M,2) = lastKFKeyPoint of the
    #        left / right branch (:231-235 / :211-215)
    #   newKeys: TrackedKeys of newKF; unMatchedF / unMatchedFR = newKF->unMatchedF / unMatchedFR
    # Returns (nMatches, out) with out[i] = (lIdx, rIdx) appended to matchedIdxs[i], or (-2, -2) if nothing was appended.
    def matchByProjectionRPredLBA(self, pts: dict, newKeys: TrackedKeys, unMatchedF, unMatchedFR, rad):
        matchDistLBA, ratioLBA = 50, _F32(0.6)
        rad = _F32(rad)
        M = len(pts["desc"])
        out = np.full((M, 2), -2, np.int32)
        nMatches = 0
        for i in range(M):
            mpDesc = pts["desc"][i]
            predScale = int(pts["predScale"][i])
            bestDist, bestIdx, bestLev, bestLev2, secDist = 256, -1, -1, -1, 256
            pLeft = pts["predL"][i]
            radius = _F32(self.feLeft.scalePyramid[predScale] * rad)
            if pLeft[0] > 0 and pLeft[1] > 0:
                for idx in self.getMatchIdxs(pLeft, newKeys, predScale, radius, False):
                    if unMatchedF[idx] >= 0:
                        continue
                    lev = int(newKeys.keyPoints["octave"][idx])
                    dist = DescriptorDistance(mpDesc, newKeys.Desc[idx])
                    if dist < bestDist:
                        secDist, bestLev2 = bestDist, bestLev
                        bestDist, bestLev, bestIdx = dist, lev, idx
                        continue
                    if dist < secDist:
                        secDist, bestLev2 = dist, bestLev   # sic (:132-136): the best key's level, not this key's
            pRight = pts["predR"][i]
            bestDistR, bestIdxR, bestLevR, bestLevR2, secDistR = 256, -1, -1, -1, 256
            if pRight[0] > 0 and pRight[1] > 0:
                for idx in self.getMatchIdxs(pRight, newKeys, predScale, radius, True):
                    if unMatchedFR[idx] >= 0:
                        continue
                    lev = int(newKeys.rightKeyPoints["octave"][idx])
                    dist = DescriptorDistance(mpDesc, newKeys.rightDesc[idx])
                    if dist < bestDistR:
                        secDistR, bestLevR2 = bestDistR, bestLevR
                        bestDistR, bestLevR, bestIdxR = dist, lev, idx
                        continue
                    if dist < secDistR:
                        secDistR, bestLevR2 = dist, lev
            right = False
            if bestDist > bestDistR:
                bestDist, secDist, bestLev, bestLev2, right = bestDistR, secDistR, bestLevR, bestLevR2, True
            if bestDist > matchDistLBA:
                continue
            thr = _F32(ratioLBA * _F32(secDist))
            if bestLev == bestLev2 and _F32(bestDist) >= thr:
                continue
            if bestLev != bestLev2 or _F32(bestDist) < thr:
                last = pts["lastR"][i] if right else pts["lastL"][i]
                pred = pRight if right else pLeft
                dx, dy = float(pred[0]) - float(last[0]), float(pred[1]) - float(last[1])
                if not (math.sqrt(dx * dx + dy * dy) > 10.0):   # Converter::checkPixelParallax, include/Conversions.h:140-144
                    continue
                if right:
                    out[i] = (newKeys.leftIdxs[bestIdxR] if newKeys.leftIdxs[bestIdxR] >= 0 else -1, bestIdxR)
                else:
                    out[i] = (bestIdx, newKeys.rightIdxs[bestIdx] if newKeys.rightIdxs[bestIdx] >= 0 else -1)
                nMatches += 1
        return nMatches, out


def prepareImage(image: np.ndarray, mapx: np.ndarray | None = None, mapy: np.ndarray | None = None) -> np.ndarray:
    """Upstream image preparation: cv::remap(INTER_LINEAR) with the CV_32F rectification maps
    (src/VIOSlam.cpp:282-301), then cvtColor BGR2GRAY (src/FeatureTracker.cpp:1131-1135)."""
    if mapx is not None:
        image = cv2.remap(image, mapx, mapy, cv2.INTER_LINEAR)
    if image.ndim == 3:
        image = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    return image


def worldToFrame(wPos: np.ndarray, predPoseInv: np.ndarray, fx, fy, cx, cy, width, height, maxScaleDist, logScale, nScaleLev,
                 pred: np.ndarray, scaleLevel: np.ndarray, inFrame: np.ndarray) -> np.ndarray:
    """FeatureTracker::worldToFrame (src/FeatureTracker.cpp:685-741) + MapPoint::predictScale (src/Map.cpp:13-23) for a batch of
    map points and one camera.  pred / scaleLevel are in-out (untouched when the point is not in the frame); returns the bools."""
    M = len(wPos)
    ok = np.zeros(M, bool)
    for i in range(M):
        w = [float(v) for v in wPos[i]]
        pt = [((float(predPoseInv[r, 0]) * w[0] + float(predPoseInv[r, 1]) * w[1]) + float(predPoseInv[r, 2]) * w[2]) + float(predPoseInv[r, 3]) * w[3]
              for r in range(4)]
        if pt[2] <= 0.0:
            inFrame[i] = 0
            continue
        invZ = 1.0 / pt[2]
        u = fx * pt[0] * invZ + cx
        v = fy * pt[1] * invZ + cy
        if u < 0 or v < 0 or u >= width or v >= height:
            inFrame[i] = 0
            continue
        dist = _F32(math.sqrt(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]))
        dif = _F32(_F32(maxScaleDist[i]) / dist)
        scale = int(math.ceil(math.log(float(dif)) / float(_F32(logScale[i]))))
        scale = 0 if scale < 0 else min(scale, int(nScaleLev[i]) - 1)
        scaleLevel[i] = scale
        inFrame[i] = 1
        pred[i] = (_F32(u), _F32(v))
        ok[i] = True
    return ok


def predictMapPoints(mode, toCamera, toCameraR, fx, fy, cx, cy, width, height, valid, wPos, maxScaleDist, logScale, nScaleLev,
                     predL, scaleLevelL, inFrame, predR=None, scaleLevelR=None, inFrameR=None, matches=None, matchedL=None,
                     matchedR=None, outliers=None):
    """FeatureTracker::removeOutOfFrameMPs (src/FeatureTracker.cpp:910-940; mode 0), removeOutOfFrameMPsMono (:942-967; mode 1),
    PredictMPsPosition (:969-1013; mode 2) on top of worldToFrame above; returns keep (modes 0/1)."""
    M = len(wPos)
    keep = np.zeros(M, np.uint8)
    for i in range(M):
        if valid is not None and not valid[i]:
            continue
        one = slice(i, i + 1)
        c1 = bool(worldToFrame(wPos[one], toCamera, fx, fy, cx, cy, width, height, maxScaleDist[one], logScale[one], nScaleLev[one],
                               predL[one], scaleLevelL[one], inFrame[one])[0])
        c2 = True
        if mode != 1:
            c2 = bool(worldToFrame(wPos[one], toCameraR, fx, fy, cx, cy, width, height, maxScaleDist[one], logScale[one], nScaleLev[one],
                                   predR[one], scaleLevelR[one], inFrameR[one])[0])
        if mode != 2:
            keep[i] = 1 if (c1 and c2) else 0
            continue
        if not c1 and matches[i][0] >= 0:
            matchedL[matches[i][0]] = -1
            matches[i][0] = -1
        if not c2 and matches[i][1] >= 0:
            matchedR[matches[i][1]] = -1
            matches[i][1] = -1
        if outliers[i]:
            outliers[i] = 0
            if matches[i][0] >= 0:
                matchedL[matches[i][0]] = -1
                matches[i][0] = -1
            if matches[i][1] >= 0:
                matchedR[matches[i][1]] = -1
                matches[i][1] = -1
    return keep


def _check2dError(P, w, obs, fx, fy, cx, cy, thres, weight):
    """FeatureTracker::check2dError (src/FeatureTracker.cpp:147-164); returns (outlier, z)."""
    p = [((float(P[r, 0]) * w[0] + float(P[r, 1]) * w[1]) + float(P[r, 2]) * w[2]) + float(P[r, 3]) * w[3] for r in range(3)]
    if p[2] <= 0:
        return True, p[2]
    invZ = 1.0 / p[2]
    u = fx * p[0] * invZ + cx
    v = fy * p[1] * invZ + cy
    eU = float(obs[0]) - u
    eV = float(obs[1]) - v
    return (eU * eU + eV * eV) * weight > thres, p[2]


def findOutliers(mode, estimPose, toCameraR, fx, fy, cx, cy, thres, closeDist, invSigma, valid, wPos, inFrame, inFrameR, matches,
                 outliers, keysL, keysR, depth, close, rightIdxs, leftIdxs):
    """FeatureTracker::findOutliersR (src/FeatureTracker.cpp:582-649; mode 0) / findOutliersMono (:651-684; mode 1).
    keysL / keysR: structured keypoint arrays (x, y, octave); in-out arrays are modified; returns (nInliers increments, nStereo)."""
    nIn = nStereo = 0
    closeDist = float(_F32(closeDist))
    for i in range(len(matches)):
        if valid is not None and not valid[i]:
            continue
        w = [float(v) for v in wPos[i]]
        right = False
        if matches[i][0] >= 0:
            if inFrame is not None and not inFrame[i]:
                continue
            P, nIdx = estimPose, int(matches[i][0])
            k = keysL[nIdx]
        elif mode == 0 and matches[i][1] >= 0:
            if inFrameR is not None and not inFrameR[i]:
                continue
            right, P, nIdx = True, toCameraR, int(matches[i][1])
            k = keysR[nIdx]
        else:
            continue
        outlier, z = _check2dError(P, w, (k["x"], k["y"]), fx, fy, cx, cy, thres, float(invSigma[int(k["octave"])]))
        outliers[i] = 1 if outlier else 0
        if outlier:
            continue
        nIn += 1
        if mode == 0 and z < closeDist and close[nIdx] and not right:
            if matches[i][1] < 0:
                continue
            k2 = keysR[int(matches[i][1])]
            outlierR, _ = _check2dError(toCameraR, w, (k2["x"], k2["y"]), fx, fy, cx, cy, thres, float(invSigma[int(k2["octave"])]))
            if not outlierR:
                nStereo += 1
            else:
                depth[nIdx] = -1
                close[nIdx] = 0
                rIdx = int(rightIdxs[nIdx])
                rightIdxs[nIdx] = -1
                if rIdx >= 0:
                    leftIdxs[rIdx] = -1
                matches[i][1] = -1
    return nIn, nStereo


def calcDescriptor(descs: np.ndarray) -> int:
    """MapPoint::calcDescriptor (src/Map.cpp:145-210): index of the descriptor with the least median distance to the rest."""
    N = len(descs)
    D = np.zeros((N, N), np.int32)
    for i in range(N):
        for j in range(i + 1, N):
            D[i, j] = D[j, i] = DescriptorDistance(descs[i], descs[j])
    best, bestIdx = 2 ** 31 - 1, 0
    for i in range(N):
        median = int(np.sort(D[i])[int(0.5 * (N - 1))])
        if median < best:
            best, bestIdx = median, i
    return bestIdx


# ----------------------------------------------------------------------------------------------
# FeatureTracker glue (src/FeatureTracker.cpp:28-70)
# ----------------------------------------------------------------------------------------------
def assignKeysToGrids(keysLeft: TrackedKeys, keypoints: np.ndarray, width: int, height: int):
    imageRatio = _F32(_F32(width) / _F32(height))
    keysLeft.xGrids = 64
    keysLeft.yGrids = int(math.ceil(_F32(_F32(keysLeft.xGrids) / imageRatio)))
    keysLeft.xMult = _F32(_F32(keysLeft.xGrids) / _F32(width))
    keysLeft.yMult = _F32(_F32(keysLeft.yGrids) / _F32(height))
    grid = [[[] for _ in range(keysLeft.xGrids)] for _ in range(keysLeft.yGrids)]
    for k in range(len(keypoints)):
        xPos = cv_round(_F32(keypoints["x"][k] * keysLeft.xMult))
        yPos = cv_round(_F32(keypoints["y"][k] * keysLeft.yMult))
        xPos = min(max(xPos, 0), keysLeft.xGrids - 1)
        yPos = min(max(yPos, 0), keysLeft.yGrids - 1)
        grid[yPos][xPos].append(k)
    return grid


def extractORBAndStereoMatch(feLeft: FeatureExtractor, feRight: FeatureExtractor, fm: FeatureMatcher,
                             leftIm: np.ndarray, rightIm: np.ndarray, width: int, height: int,
                             stats: dict | None = None) -> TrackedKeys:
    k = TrackedKeys()
    k.keyPoints, k.Desc = feLeft.extractKeysNew(leftIm, stats)
    k.rightKeyPoints, k.rightDesc = feRight.extractKeysNew(rightIm)
    fm.findStereoMatchesORB2R(k.rightDesc, k.rightKeyPoints, k, stats)
    k.lkeyGrid = assignKeysToGrids(k, k.keyPoints, width, height)
    k.rkeyGrid = assignKeysToGrids(k, k.rightKeyPoints, width, height)
    return k
