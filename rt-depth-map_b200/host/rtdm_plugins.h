// rtdm_plugins.h -- C++ plugin peers for rt-depth-map, header-only, over the C ABI (include/rtdm_b200.h).
//
//   CUDAMatcherKonolige      : BlockMatcher        peer of SWMatcherKonolige     (reference include/stereo-matcher/bm-sw.h:25-37)
//   CUDASemiGlobalMatcher    : BlockMatcher        peer of SWSemiGlobalMatcher   (reference include/stereo-matcher/sgbm-sw.h:25-36)
//   CUDAMorphologicalFilter  : VideoFilterDevice   peer of SWMorphologicalFilter (reference include/filter/mf-sw.h:16-21)
//
// Same constructor argument lists, same virtuals, same return conventions, so Estimator
// (estimator.cpp:45,54-56) needs no change: main.cpp:133-135 picks them with `new` exactly like the SW /
// HW variants.  Inside the reference tree, define RTDM_REFERENCE_TREE before including this file and the
// reference's own "stereo-matcher/stereo-matcher.h" / "filter/filter.h" ABCs are used; stand-alone (tests)
// the identical ABCs below are used.
//
// Differences from the SW plugins, by design: compute() runs on the GPU and returns -1 (after printing
// the library's error text, the HW plugins' convention, generic-hw-filter-ip.cpp:130-135) instead of
// throwing cv::Exception; constructors throw std::runtime_error when no device is usable (the FPGA
// plugins call exit(1), bm-hw-ip.cpp:135-168).  There is no CPU fallback.
#pragma once

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtdm_b200.h"
#include "cv_compat.h"

#ifdef RTDM_REFERENCE_TREE
#include "stereo-matcher/stereo-matcher.h"
#include "filter/filter.h"
#else
// identical to the reference's ABCs (stereo-matcher.h:13-19, filter.h:13-37)
class BlockMatcher {
public:
    virtual ~BlockMatcher() {}
    virtual int compute(cv::InputArray left, cv::InputArray right, cv::OutputArray out) = 0;
    virtual void setROI1(cv::Rect roi1) = 0;
    virtual void setROI2(cv::Rect roi2) = 0;
};

class VideoFilterDevice {
public:
    virtual ~VideoFilterDevice() {}
    int getFrameSize() const { return img_width * img_height * (img_bpp >> 3); }
    int getBpp() const { return img_bpp; }
    void setBpp(int bpp) { img_bpp = bpp; }
    int getWidth() const { return img_width; }
    void setWidth(int v) { img_width = v; }
    int getHeight() const { return img_height; }
    void setHeight(int v) { img_height = v; }
    char *getVideoInBuffer() { return video_in; }
    char *getVideoOutBuffer() { return video_out; }
    virtual int run(cv::InputArray in, cv::OutputArray out) = 0;
protected:
    int img_width = 0, img_height = 0, img_bpp = 0;
    char *video_in = nullptr;
    char *video_out = nullptr;
    const char *dev_name = nullptr;
    int dev_minor = 0;
};
#endif

namespace rtdm_detail {
inline void fail(const char *what, int rc)
{
    throw std::runtime_error(std::string(what) + ": rtdm error " + std::to_string(rc) + ": " + rtdm_last_error());
}
}  // namespace rtdm_detail

class CUDAMatcherKonolige : public BlockMatcher {
public:
    // same arguments as SWMatcherKonolige (bm-sw.h:28-30); roi1, roi2 and maxDisparity are accepted and
    // unused exactly as in the reference (bm-sw.cpp:12-14 vs :16-25).  max_width/max_height bound the frames.
    CUDAMatcherKonolige(cv::Rect &roi1, cv::Rect &roi2, int preFilterCap, int blockSize, int minDisparity,
                        int textureThreshold, int numOfDisparities, int maxDisparity, int uniquenessRatio,
                        int speckleWindowSize, int speckleRange, int disp12MaxDiff,
                        int max_width = 1280, int max_height = 720, int device = 0)
    {
        (void)roi1; (void)roi2; (void)maxDisparity;
        rtdm_params p;
        rtdm_params_default_bm(&p);
        p.preFilterCap = preFilterCap; p.blockSize = blockSize; p.minDisparity = minDisparity;
        p.textureThreshold = textureThreshold; p.numDisparities = numOfDisparities;
        p.uniquenessRatio = uniquenessRatio; p.speckleWindowSize = speckleWindowSize;
        p.speckleRange = speckleRange; p.disp12MaxDiff = disp12MaxDiff;
        int rc = rtdm_bm_create(&h_, &p, max_width, max_height, 1, device);
        if (rc) rtdm_detail::fail("CUDAMatcherKonolige", rc);
    }
    ~CUDAMatcherKonolige() { rtdm_bm_destroy(h_); }
    void setROI1(cv::Rect r) override { rtdm_bm_set_roi1(h_, r.x, r.y, r.width, r.height); }
    void setROI2(cv::Rect r) override { rtdm_bm_set_roi2(h_, r.x, r.y, r.width, r.height); }
    int compute(cv::InputArray left, cv::InputArray right, cv::OutputArray out) override
    {
        const cv::Mat l = left.getMat(), r = right.getMat();
        out.create(l.rows, l.cols, CV_16SC1);
        cv::Mat d = out.getMat();
        int rc = rtdm_bm_compute(h_, l.data, l.step, r.data, r.step, l.cols, l.rows, (int16_t *)d.data, d.step);
        if (rc) { std::fprintf(stderr, "CUDAMatcherKonolige::compute: %s\n", rtdm_last_error()); return -1; }
        return 0;
    }
private:
    rtdm_bm *h_ = nullptr;
};

// One large frame as row bands over several GPUs of this process (rtdm_bm_rowband_*: every device gets its own input
// rows + halo, the int16 bands come back to devices[0] as peer copies, the speckle filter runs there).  Same constructor
// arguments as SWMatcherKonolige plus the device list; bit-exact against CUDAMatcherKonolige.  minDisparity <= 0.
class CUDARowBandMatcherKonolige : public BlockMatcher {
public:
    CUDARowBandMatcherKonolige(cv::Rect &roi1, cv::Rect &roi2, int preFilterCap, int blockSize, int minDisparity,
                               int textureThreshold, int numOfDisparities, int maxDisparity, int uniquenessRatio,
                               int speckleWindowSize, int speckleRange, int disp12MaxDiff,
                               int n_gpus, const int *devices, int max_width = 1280, int max_height = 720)
    {
        (void)roi1; (void)roi2; (void)maxDisparity;
        rtdm_params p;
        rtdm_params_default_bm(&p);
        p.preFilterCap = preFilterCap; p.blockSize = blockSize; p.minDisparity = minDisparity;
        p.textureThreshold = textureThreshold; p.numDisparities = numOfDisparities;
        p.uniquenessRatio = uniquenessRatio; p.speckleWindowSize = speckleWindowSize;
        p.speckleRange = speckleRange; p.disp12MaxDiff = disp12MaxDiff;
        int rc = rtdm_bm_rowband_create(&h_, &p, max_width, max_height, n_gpus, devices);
        if (rc) rtdm_detail::fail("CUDARowBandMatcherKonolige", rc);
    }
    ~CUDARowBandMatcherKonolige() { rtdm_bm_rowband_destroy(h_); }
    void setROI1(cv::Rect r) override { rtdm_bm_rowband_set_roi1(h_, r.x, r.y, r.width, r.height); }
    void setROI2(cv::Rect r) override { rtdm_bm_rowband_set_roi2(h_, r.x, r.y, r.width, r.height); }
    int compute(cv::InputArray left, cv::InputArray right, cv::OutputArray out) override
    {
        const cv::Mat l = left.getMat(), r = right.getMat();
        out.create(l.rows, l.cols, CV_16SC1);
        cv::Mat d = out.getMat();
        int rc = rtdm_bm_rowband_compute(h_, l.data, l.step, r.data, r.step, l.cols, l.rows, (int16_t *)d.data, d.step);
        if (rc) { std::fprintf(stderr, "CUDARowBandMatcherKonolige::compute: %s\n", rtdm_last_error()); return -1; }
        return 0;
    }
private:
    rtdm_bm_rowband *h_ = nullptr;
};

class CUDASemiGlobalMatcher : public BlockMatcher {
public:
    // same arguments as SWSemiGlobalMatcher (sgbm-sw.h:28-29); P1 = 8*3*5*5, P2 = 32*3*5*5 as in
    // sgbm-sw.cpp:17-18.  `mode` keeps the reference's default MODE_SGBM; RTDM_SGBM_MODE_HH selects 8 paths.
    CUDASemiGlobalMatcher(int blockSize, int minDisparity, int numOfDisparities, int uniquenessRatio,
                          int speckleWindowSize, int speckleRange, int disp12MaxDiff,
                          int mode = RTDM_SGBM_MODE_SGBM, int max_width = 1280, int max_height = 720, int device = 0)
    {
        rtdm_params p;
        rtdm_params_default_sgbm(&p);
        p.blockSize = blockSize; p.minDisparity = minDisparity; p.numDisparities = numOfDisparities;
        p.uniquenessRatio = uniquenessRatio; p.speckleWindowSize = speckleWindowSize;
        p.speckleRange = speckleRange; p.disp12MaxDiff = disp12MaxDiff; p.mode = mode;
        int rc = rtdm_sgbm_create(&h_, &p, max_width, max_height, 1, device);
        if (rc) rtdm_detail::fail("CUDASemiGlobalMatcher", rc);
    }
    ~CUDASemiGlobalMatcher() { rtdm_sgbm_destroy(h_); }
    void setROI1(cv::Rect) override {}       // no-ops like the reference (sgbm-sw.h:32-33)
    void setROI2(cv::Rect) override {}
    int compute(cv::InputArray left, cv::InputArray right, cv::OutputArray out) override
    {
        const cv::Mat l = left.getMat(), r = right.getMat();
        out.create(l.rows, l.cols, CV_16SC1);
        cv::Mat d = out.getMat();
        int rc = rtdm_sgbm_compute(h_, l.data, l.step, r.data, r.step, l.cols, l.rows, (int16_t *)d.data, d.step);
        if (rc) { std::fprintf(stderr, "CUDASemiGlobalMatcher::compute: %s\n", rtdm_last_error()); return -1; }
        return 0;
    }
private:
    rtdm_sgbm *h_ = nullptr;
};

class CUDAMorphologicalFilter : public VideoFilterDevice {
public:
    // same arguments as SWMorphologicalFilter(w, h, bpp) (mf-sw.cpp:10-17).  video_in / video_out are the
    // handle's PINNED host buffers: Estimator writes the inRange() mask into video_in (estimator.cpp:43)
    // and wraps both in cv::Mat headers (estimator.cpp:141-142), so the copies to/from the GPU are DMA.
    explicit CUDAMorphologicalFilter(int w, int h, int bpp, int device = 0)
    {
        int rc = rtdm_morph_create(&h_, w, h, bpp, 1, device);
        if (rc) rtdm_detail::fail("CUDAMorphologicalFilter", rc);
        img_width = w; img_height = h; img_bpp = bpp;
        video_in = (char *)rtdm_morph_in_buffer(h_);
        video_out = (char *)rtdm_morph_out_buffer(h_);
    }
    ~CUDAMorphologicalFilter() { rtdm_morph_destroy(h_); }
    // erode, dilate, dilate, erode with the 10x10 ellipse (mf-sw.cpp:22-27).  Returns 0 (the reference's
    // run() has no return statement, mf-sw.cpp:19-28), -1 on failure like the HW filter.
    int run(cv::InputArray in, cv::OutputArray out) override
    {
        const cv::Mat i = in.getMat();
        out.create(i.rows, i.cols, CV_8UC1);
        cv::Mat o = out.getMat();
        if (i.cols != img_width || i.rows != img_height || i.step != (size_t)img_width || o.step != (size_t)img_width) {
            std::fprintf(stderr, "CUDAMorphologicalFilter::run: frames must be %dx%d, tightly packed\n", img_width, img_height);
            return -1;
        }
        int rc = rtdm_morph_run(h_, i.data, o.data);
        if (rc) { std::fprintf(stderr, "CUDAMorphologicalFilter::run: %s\n", rtdm_last_error()); return -1; }
        return 0;
    }
private:
    rtdm_morph *h_ = nullptr;
};

// The step before the matcher, fused on the GPU.  Replaces in Estimator::run, per camera,
//     cvtColor(img[i], gray, CV_RGB2GRAY); remap(gray, rect, map1, map2, INTER_LINEAR); rect = rect(roif);   estimator.cpp:29-36
// by   rectifier.run(img[i], rect);      // rect: CV_8UC1, roif.height x roif.width
// map1 / map2 are the CV_16SC2 / CV_16UC1 maps main.cpp:95-96 builds with initUndistortRectifyMap.
class CUDARectifier {
public:
    CUDARectifier(const cv::Mat &map1, const cv::Mat &map2, const cv::Rect &roif, int device = 0) : roi_(roif)
    {
        int rc = rtdm_rectify_create(&h_, map2.cols, map2.rows, map1.ptr<short>(), map1.step, map2.ptr<unsigned short>(), map2.step,
                                     roif.x, roif.y, roif.width, roif.height, 1, device);
        if (rc) rtdm_detail::fail("CUDARectifier", rc);
    }
    ~CUDARectifier() { rtdm_rectify_destroy(h_); }
    // rgb: CV_8UC3 frame in R,G,B order (the decoder's output, estimator.cpp:24-27); rect is created like OutputArray::create
    int run(const cv::Mat &rgb, cv::Mat &rect)
    {
        rect.create(roi_.height, roi_.width, CV_8UC1);
        int rc = rtdm_rectify_run(h_, 1, rgb.ptr<unsigned char>(), rgb.step, 0, rect.ptr<unsigned char>(), rect.step, 0);
        if (rc) { std::fprintf(stderr, "CUDARectifier::run: %s\n", rtdm_last_error()); return -1; }
        return 0;
    }
private:
    rtdm_rectify *h_ = nullptr;
    cv::Rect roi_;
};

// The step before the morphological filter, fused on the GPU.  Replaces in Estimator::run
//     remap(img[0], img_rectified, remap_left1, remap_left2, INTER_LINEAR); img_rectified = img_rectified(roif);   estimator.cpp:38-39
//     cvtColor(img_rectified, img_rectified, COLOR_RGB2BGR); cvtColor(img_rectified, imgHSV, COLOR_BGR2HSV);      :40,:42
//     inRange(imgHSV, Scalar(iLowH, iLowS, iLowV), Scalar(iHighH, iHighS, iHighV), filter_in);                    :43
// by   colormask.run(img[0], low, high, filter_in, &img_rectified);
// filter_in may wrap the filter plugin's own input buffer, as in the reference (estimator.cpp:141).
class CUDAColorMask {
public:
    CUDAColorMask(const cv::Mat &map1, const cv::Mat &map2, const cv::Rect &roif, int device = 0) : roi_(roif)
    {
        int rc = rtdm_colormask_create(&h_, map2.cols, map2.rows, map1.ptr<short>(), map1.step, map2.ptr<unsigned short>(), map2.step,
                                       roif.x, roif.y, roif.width, roif.height, 1, device);
        if (rc) rtdm_detail::fail("CUDAColorMask", rc);
    }
    ~CUDAColorMask() { rtdm_colormask_destroy(h_); }
    // rgb: CV_8UC3 frame in R,G,B order; low / high: (H, S, V); filter_in: CV_8UC1 roif.height x roif.width (created if
    // it has another size); bgr_rectified: optional CV_8UC3 buffer of roif.height rows with step >= 3 * roif.width
    int run(const cv::Mat &rgb, const int low[3], const int high[3], cv::Mat &filter_in, cv::Mat *bgr_rectified = nullptr)
    {
        if (filter_in.rows != roi_.height || filter_in.cols != roi_.width) filter_in.create(roi_.height, roi_.width, CV_8UC1);
        int rc = rtdm_colormask_run(h_, 1, rgb.ptr<unsigned char>(), rgb.step, 0, low, high, filter_in.ptr<unsigned char>(), filter_in.step, 0,
                                    bgr_rectified ? bgr_rectified->ptr<unsigned char>() : nullptr, bgr_rectified ? bgr_rectified->step : 0, 0);
        if (rc) { std::fprintf(stderr, "CUDAColorMask::run: %s\n", rtdm_last_error()); return -1; }
        return 0;
    }
private:
    rtdm_colormask *h_ = nullptr;
    cv::Rect roi_;
};

// The step after the morphological filter on the GPU.  Replaces in Estimator::run
//     filter_out.copyTo(contInput); findContours(contInput, contours, hierarchy, CV_RETR_EXTERNAL, CV_CHAIN_APPROX_SIMPLE);   estimator.cpp:46-47
//     fill_bounding_rects_of_contours(contours, hierarchy, obj_boundings, minObjSize);                                        :51 (:164-175)
//     find_relevant_matching_region(obj_boundings, matching_roi);                                                             :52 (:177-204)
// by   if (regions.run(filter_out, minObjSize, obj_boundings, matching_roi) > 0) { bm->setROI1(matching_roi); ... }
// run returns contours.size() (the reference skips the matcher when it is 0), or -1 on error.
class CUDAObjectRegions {
public:
    CUDAObjectRegions(int max_width, int max_height, int max_regions = 256, int device = 0) : max_(max_regions), buf_(4 * (size_t)max_regions)
    {
        int rc = rtdm_regions_create(&h_, max_width, max_height, max_regions, device);
        if (rc) rtdm_detail::fail("CUDAObjectRegions", rc);
    }
    ~CUDAObjectRegions() { rtdm_regions_destroy(h_); }
    int run(const cv::Mat &filter_out, int min_obj_size, std::vector<cv::Rect> &obj_boundings, cv::Rect &matching_roi)
    {
        int n = 0, nc = 0, roi[4] = {0, 0, 0, 0};
        int rc = rtdm_regions_run(h_, filter_out.ptr<unsigned char>(), filter_out.step, filter_out.cols, filter_out.rows, min_obj_size,
                                  buf_.data(), &n, &nc, roi);
        if (rc) { std::fprintf(stderr, "CUDAObjectRegions::run: %s\n", rtdm_last_error()); return -1; }
        obj_boundings.clear();
        for (int i = 0; i < n; i++) obj_boundings.push_back(cv::Rect(buf_[4 * i], buf_[4 * i + 1], buf_[4 * i + 2], buf_[4 * i + 3]));
        matching_roi = cv::Rect(roi[0], roi[1], roi[2], roi[3]);
        return nc;
    }
private:
    rtdm_regions *h_ = nullptr;
    int max_;
    std::vector<int> buf_;
};

// The step after the matcher, fused on the GPU.  Replaces in Estimator::run
//     left_disp /= 16.;                                              estimator.cpp:75
//     reprojectImageTo3D(left_disp, xyz, Q, true, CV_32F);           estimator.cpp:76
//     calc_depth(xyz, left_disp, filter_out, img_rectified, obj_boundings, calibration_unit);   :77
// by   depth.run(left_disp, Q, filter_out, obj_boundings, mean_z);   // left_disp still x16, as the matcher wrote it
// and the label of region i is  mean_z[i] * calibration_unit / 10  cm  (estimator.cpp:252-254), drawn when count[i] > 0.
class CUDADepthEpilogue {
public:
    CUDADepthEpilogue(int max_width, int max_height, int max_regions = 64, int device = 0)
    {
        int rc = rtdm_depth_create(&h_, max_width, max_height, max_regions, device);
        if (rc) rtdm_detail::fail("CUDADepthEpilogue", rc);
    }
    ~CUDADepthEpilogue() { rtdm_depth_destroy(h_); }
    // disp: CV_16SC1 x16 disparity; Q: 4x4 doubles, row-major (cv::Mat_<double>(4,4).ptr<double>()); mask: CV_8UC1 or empty
    int run(const cv::Mat &disp, const double *Q, const cv::Mat &mask, const std::vector<cv::Rect> &regions,
            std::vector<double> &mean_z, std::vector<int> *count = nullptr)
    {
        const int n = (int)regions.size();
        std::vector<int> rects((size_t)4 * n), cnt((size_t)n);
        for (int i = 0; i < n; i++) { rects[4 * i] = regions[i].x; rects[4 * i + 1] = regions[i].y; rects[4 * i + 2] = regions[i].width; rects[4 * i + 3] = regions[i].height; }
        mean_z.assign((size_t)n, 0.0);
        int rc = rtdm_depth_run(h_, disp.ptr<short>(), disp.step, disp.cols, disp.rows, Q, mask.empty() ? nullptr : mask.ptr<unsigned char>(),
                                mask.empty() ? 0 : mask.step, n, rects.data(), mean_z.data(), cnt.data(), nullptr, 0);
        if (rc) { std::fprintf(stderr, "CUDADepthEpilogue::run: %s\n", rtdm_last_error()); return -1; }
        if (count) *count = cnt;
        return 0;
    }
private:
    rtdm_depth *h_ = nullptr;
};
