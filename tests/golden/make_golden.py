"""Generates the committed golden fixtures with cv2 4.13.0 (the in-image build of the OpenCV
routines the reference's SW plugins call).  Run from the repo root:  python tests/golden/make_golden.py

Each .npz holds the inputs, the parameters (as a JSON string) and the cv2 output, so the fixtures
are self-contained: neither cv2 nor /root/reference is needed to check against them.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rt-depth-map_b200"))
from oracle import cv2_ref  # noqa: E402
from rtdm_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

BM_CASES = {
    # name: (W, H, seed, params)
    "bm_cfg0_320x240_nd64_bs15": (320, 240, 1000, dict(preFilterCap=31, blockSize=15, minDisparity=0, textureThreshold=10, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)),
    "bm_cfg1_640x480_nd128_bs13": (640, 480, 1001, dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=128, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)),
    "bm_raw_240x160_nd32_bs9": (240, 160, 1002, dict(preFilterCap=31, blockSize=9, minDisparity=0, textureThreshold=10, numDisparities=32, uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)),
    "bm_oddh_233x157_nd48_bs21": (233, 157, 1003, dict(preFilterCap=31, blockSize=21, minDisparity=0, textureThreshold=10, numDisparities=48, uniquenessRatio=0, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)),
    "bm_roi_320x240_nd64_bs15": (320, 240, 1004, dict(preFilterCap=31, blockSize=15, minDisparity=0, textureThreshold=10, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, roi1=(40, 30, 240, 180), roi2=(10, 20, 280, 200))),
    "bm_norm_320x240_nd64_bs11": (320, 240, 1005, dict(preFilterCap=25, blockSize=11, minDisparity=0, textureThreshold=20, numDisparities=64, uniquenessRatio=15, speckleWindowSize=50, speckleRange=16, disp12MaxDiff=2, preFilterType=0, preFilterSize=9)),
    "bm_cap63_200x120_nd48_bs11": (200, 120, 1006, dict(preFilterCap=63, blockSize=11, minDisparity=0, textureThreshold=10, numDisparities=48, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=-1)),
    # minDisparity != 0 (a constructor argument of both reference peers, bm-sw.h:28-30): positive values make cv2 write the
    # last minD computed columns of a row into the next row -- the ROI case keeps such a row visible
    "bm_mind16_320x240_nd64_bs15": (320, 240, 1007, dict(preFilterCap=31, blockSize=15, minDisparity=16, textureThreshold=10, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)),
    "bm_mindneg16_320x240_nd64_bs13": (320, 240, 1008, dict(preFilterCap=31, blockSize=13, minDisparity=-16, textureThreshold=10, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)),
    "bm_mind16roi_320x240_nd64_bs15": (320, 240, 99, dict(preFilterCap=31, blockSize=15, minDisparity=16, textureThreshold=10, numDisparities=64, uniquenessRatio=10, speckleWindowSize=0, speckleRange=32, disp12MaxDiff=-1, roi1=(30, 20, 250, 180))),
    # the reference's real operating point at 1280x720: the calibrated ROI crop 934x404 (backup/1280x720/extrinsics.yml:56-57 via
    # main.cpp:80-85), -nd 192 (cmdline-parser.cpp:22), main.cpp:134-135's literals, setROI1 from the object boxes (estimator.cpp:54)
    "bm_op_934x404_nd192_bs13": (934, 404, 1009, dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=192, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, roi1=(260, 60, 520, 280))),
}

SGBM_CASES = {
    "sgbm_mode0_320x240_nd64_bs5": (320, 240, 2000, dict(blockSize=5, minDisparity=0, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=0)),
    "sgbm_hh_320x240_nd64_bs5": (320, 240, 2001, dict(blockSize=5, minDisparity=0, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=1)),
    "sgbm_mode0_233x157_nd32_bs3": (233, 157, 2002, dict(blockSize=3, minDisparity=0, numDisparities=32, uniquenessRatio=5, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=2, mode=0)),
    "sgbm_hh_233x157_nd48_bs7": (233, 157, 2003, dict(blockSize=7, minDisparity=0, numDisparities=48, uniquenessRatio=15, speckleWindowSize=60, speckleRange=8, disp12MaxDiff=1, mode=1)),
    # minDisparity != 0 (sgbm-sw.h:28-29).  For minD >= 2 cv2's LR check treats never-written disp2 entries as disparities
    # (they hold the x16-scaled invalid value, which passes its `>= minD` test)
    "sgbm_hh_mind16_320x240_nd64_bs5": (320, 240, 2004, dict(blockSize=5, minDisparity=16, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=1)),
    "sgbm_mode0_mind16_320x240_nd64_bs5": (320, 240, 2005, dict(blockSize=5, minDisparity=16, numDisparities=64, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=0)),
    "sgbm_hh_mindneg16_233x157_nd48_bs5": (233, 157, 2006, dict(blockSize=5, minDisparity=-16, numDisparities=48, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=1)),
}


def main(only=None):
    """only: names of BM / SGBM cases to (re)generate; None = every fixture."""
    cv2 = cv2_ref.cv2_pinned()
    for name, (W, H, seed, p) in BM_CASES.items():
        if only is not None and name not in only:
            continue
        L, R, _ = synth.stereo_pair(W, H, p["numDisparities"], seed)
        disp = cv2_ref.make_bm(**p).compute(L, R)
        np.savez_compressed(os.path.join(OUT, name), left=L, right=R, disp=disp, params=json.dumps(p))
        print(name, disp.shape, float((disp >= 0).mean()))
    for name, (W, H, seed, p) in SGBM_CASES.items():
        if only is not None and name not in only:
            continue
        L, R, _ = synth.stereo_pair(W, H, p["numDisparities"], seed)
        disp = cv2_ref.make_sgbm(**p).compute(L, R)
        np.savez_compressed(os.path.join(OUT, name), left=L, right=R, disp=disp, params=json.dumps(p))
        print(name, disp.shape, float((disp >= 0).mean()))
    if only is not None:
        return
    # morphology: the ROI-sized binary mask of the 720p calibration, a full 720p gray image, tiny edge cases
    for name, img in {
        "morph_mask_934x404": synth.binary_mask(934, 404, 3000),
        "morph_gray_320x240": synth.gray_image(320, 240, 3001),
        "morph_mask_17x13": synth.binary_mask(17, 13, 3002),
        "morph_mask_10x10": synth.binary_mask(10, 10, 3003),
    }.items():
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (10, 10))
        np.savez_compressed(os.path.join(OUT, name), src=img, erode=cv2.erode(img, k), dilate=cv2.dilate(img, k),
                            openclose=cv2_ref.morph_open_close(img), se=k)
        print(name, img.shape)
    # speckle / median / validate on a raw BM disparity map
    L, R, _ = synth.stereo_pair(320, 240, 64, 4000)
    raw = cv2_ref.make_bm(31, 15, 0, 10, 64, 10, 0, 0, -1).compute(L, R)
    sp = {}
    for (ms, md) in [(100, 32), (10, 0), (1000, 16), (1, 1)]:
        t = raw.copy(); cv2.filterSpeckles(t, -16, ms, md); sp[f"sp_{ms}_{md}"] = t
    np.savez_compressed(os.path.join(OUT, "post_speckle_320x240"), raw=raw, **sp)
    rng = np.random.default_rng(4001)
    a = rng.integers(-16, 2048, (97, 131)).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "post_median_131x97"), src=a, median=cv2.medianBlur(a, 3))
    cost = rng.integers(0, 5000, raw.shape).astype(np.int16)
    v = {}
    for d12 in (0, 1, 3):
        t = raw.copy(); cv2.validateDisparity(t, cost, 0, 64, d12); v[f"d12_{d12}"] = t
    np.savez_compressed(os.path.join(OUT, "post_validate_320x240"), raw=raw, cost=cost, **v)
    # depth epilogue: /16, reprojectImageTo3D(handleMissingValues, CV_32F), calc_depth
    from oracle import oracle as orc
    L, R, _ = synth.stereo_pair(320, 240, 64, 4100)
    disp = cv2_ref.make_bm(31, 15, 0, 10, 64, 10, 100, 32, 1).compute(L, R)
    d16 = disp.copy().astype(np.int16)
    d16m = (cv2.divide(d16, 16.0)).astype(np.int16)                 # what `left_disp /= 16.` leaves in the CV_16S Mat
    # Q as cv::stereoRectify builds it (backup/1280x720/extrinsics.yml layout): [1 0 0 -cx; 0 1 0 -cy; 0 0 0 f; 0 0 -1/Tx (cx-cx')/Tx]
    Q = np.array([[1, 0, 0, -161.37], [0, 1, 0, -118.21], [0, 0, 0, 351.933], [0, 0, 1 / 59.87, -(161.37 - 160.9) / 59.87]], np.float64)
    xyz = cv2.reprojectImageTo3D(d16m, Q, handleMissingValues=True, ddepth=cv2.CV_32F)
    mask = synth.binary_mask(320, 240, 4101)
    rects = np.array([[20, 30, 100, 80], [150, 10, 160, 200], [0, 0, 320, 240], [300, 200, 20, 40], [90, 90, 1, 1]], np.int32)
    means, counts = orc.calc_depth(xyz, mask, rects)                # the reference's loop, restated (no cv2 routine to call)
    np.savez_compressed(os.path.join(OUT, "depth_320x240"), disp=disp, div16=d16m, Q=Q, xyz=xyz, mask=mask, rects=rects,
                        mean_z=means, count=counts)
    print("depth_320x240", means, counts)
    # rectification front-end: cvtColor RGB2GRAY, remap INTER_LINEAR with CV_16SC2 maps, ROI crop
    rng = np.random.default_rng(4200)
    H, W = 240, 320
    base = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3)).astype(np.uint8), (0, 0), 1.5)
    rgb = np.clip(base.astype(np.int16) + rng.integers(-20, 21, (H, W, 3)), 0, 255).astype(np.uint8)
    K = np.array([[300., 0, 160.3], [0, 301., 119.6], [0, 0, 1]]); D = np.array([-0.31, 0.12, 0.001, -0.0007, 0.0])
    Rm = cv2.Rodrigues(np.array([0.02, -0.03, 0.01]))[0]; P = np.array([[250., 0, 150, 0], [0, 250, 125, 0], [0, 0, 1, 0]])
    m1, m2 = cv2.initUndistortRectifyMap(K, D, Rm, P, (W, H), cv2.CV_16SC2)       # main.cpp:95-96
    gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)                                   # estimator.cpp:29
    rect = cv2.remap(gray, m1, m2, cv2.INTER_LINEAR)                               # estimator.cpp:32
    roi = (21, 13, 270, 200)
    crop = np.ascontiguousarray(rect[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]])   # estimator.cpp:33
    np.savez_compressed(os.path.join(OUT, "rectify_320x240"), rgb=rgb, map1=m1, map2=m2, gray=gray, rect=rect,
                        roi=np.array(roi, np.int32), crop=crop)
    print("rectify_320x240")
    make_mask()
    print("done")


def cv_boxes(mask):
    """estimator.cpp:47 + fill_bounding_rects_of_contours (:164-175) without the size filter: boundingRect of the
    top-level contours, walking hierarchy[i][0] from contour 0 like the reference."""
    import cv2
    cs, hier = cv2.findContours(mask.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    out = []
    if len(cs):
        i = 0
        while i >= 0:
            out.append(tuple(int(v) for v in cv2.boundingRect(cs[i]))); i = int(hier[0][i][0])
    return out


def make_mask():
    """mask front-end / back-end: estimator.cpp:38-53 with cv2 itself."""
    import cv2
    rng = np.random.default_rng(4300)
    H, W = 240, 320
    # coloured blobs on a noisy background, so that the HSV threshold selects connected objects with holes
    base = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3)).astype(np.float32), (0, 0), 9.0)
    base = (base - base.mean()) / base.std() * 70 + 128
    rgb = np.clip(base + rng.integers(-6, 7, (H, W, 3)), 0, 255).astype(np.uint8)
    K = np.array([[300., 0, 160.3], [0, 301., 119.6], [0, 0, 1]]); D = np.array([-0.31, 0.12, 0.001, -0.0007, 0.0])
    Rm = cv2.Rodrigues(np.array([0.02, -0.03, 0.01]))[0]; P = np.array([[250., 0, 150, 0], [0, 250, 125, 0], [0, 0, 1, 0]])
    m1, m2 = cv2.initUndistortRectifyMap(K, D, Rm, P, (W, H), cv2.CV_16SC2)
    roi = (21, 13, 270, 200)
    rect = cv2.remap(rgb, m1, m2, cv2.INTER_LINEAR)                                  # estimator.cpp:38
    rect = np.ascontiguousarray(rect[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]])   # :39
    bgr = cv2.cvtColor(rect, cv2.COLOR_RGB2BGR)                                      # :40
    hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)                                       # :42
    low, high = (30, 60, 50), (100, 255, 255)
    filter_in = cv2.inRange(hsv, low, high)                                          # :43
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (10, 10))
    filter_out = cv2.erode(cv2.dilate(cv2.dilate(cv2.erode(filter_in, k), k), k), k)  # mf-sw.cpp:22-27
    boxes_in, boxes_out = cv_boxes(filter_in), cv_boxes(filter_out)
    np.savez_compressed(os.path.join(OUT, "mask_320x240"), rgb=rgb, map1=m1, map2=m2, roi=np.array(roi, np.int32),
                        low=np.array(low, np.int32), high=np.array(high, np.int32), bgr=bgr, hsv=hsv, filter_in=filter_in,
                        filter_out=filter_out, boxes_in=np.array(boxes_in, np.int32).reshape(-1, 4),
                        boxes_out=np.array(boxes_out, np.int32).reshape(-1, 4))
    print("mask_320x240", float((filter_in != 0).mean()), len(boxes_in), len(boxes_out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mask":
        make_mask()
    elif len(sys.argv) > 1:
        main(set(sys.argv[1:]))         # python tests/golden/make_golden.py <case name> ...
    else:
        main()
