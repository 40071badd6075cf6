// Drives the C++ plugin peers the way Estimator does (estimator.cpp:45,54-56,141-142).
//   host_adapter_main probe                      -> prints "nodevice" (exit 3) or "ok" (exit 0)
//   host_adapter_main bm   W H nd bs left right out [rx ry rw rh]
//   host_adapter_main sgbm W H nd bs left right out mode
//   host_adapter_main morph W H in out
//   host_adapter_main rectify W H rgb map1 map2 out
//   host_adapter_main mask W H rgb map1 map2 mask_out boxes_out
//   host_adapter_main depth W H disp mask out x y w h   (Q is the fixed matrix of tests/test_host_adapters.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rtdm_plugins.h"

static std::vector<unsigned char> slurp(const char *path, size_t n)
{
    std::vector<unsigned char> b(n);
    FILE *f = std::fopen(path, "rb");
    if (!f || std::fread(b.data(), 1, n, f) != n) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
    std::fclose(f);
    return b;
}

static void dump(const char *path, const void *p, size_t n)
{
    FILE *f = std::fopen(path, "wb");
    if (!f || std::fwrite(p, 1, n, f) != n) { std::fprintf(stderr, "cannot write %s\n", path); std::exit(2); }
    std::fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    try {
        if (!std::strcmp(argv[1], "probe")) {
            cv::Rect r;
            CUDAMatcherKonolige m(r, r, 31, 13, 0, 10, 128, 128, 10, 100, 32, 1);
            std::puts("ok");
            return 0;
        }
        if (!std::strcmp(argv[1], "bm") && argc >= 9) {
            int W = atoi(argv[2]), H = atoi(argv[3]), nd = atoi(argv[4]), bs = atoi(argv[5]);
            // a frame embedded in a wider buffer: the matcher sees a non-contiguous ROI view (estimator.cpp:33,36)
            const int PAD = 24;
            std::vector<unsigned char> l = slurp(argv[6], (size_t)W * H), r = slurp(argv[7], (size_t)W * H);
            cv::Mat lf(H, W + PAD, CV_8UC1), rf(H, W + PAD, CV_8UC1);
            for (int y = 0; y < H; y++) {
                std::memcpy(lf.ptr<unsigned char>(y) + 8, &l[(size_t)y * W], W);
                std::memcpy(rf.ptr<unsigned char>(y) + 8, &r[(size_t)y * W], W);
            }
            cv::Mat lv = lf(cv::Rect(8, 0, W, H)), rv = rf(cv::Rect(8, 0, W, H));
            cv::Rect roi;
            // RTDM_TEST_ROWBANDS=k: the same frame as k row bands (all on device 0 here: bands then run one after the other)
            const char *rbenv = std::getenv("RTDM_TEST_ROWBANDS");
            const int nb = rbenv ? atoi(rbenv) : 0;
            const int devs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            BlockMatcher *bm = nb > 1 ? static_cast<BlockMatcher *>(new CUDARowBandMatcherKonolige(roi, roi, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1, nb > 8 ? 8 : nb, devs, W, H))
                                      : static_cast<BlockMatcher *>(new CUDAMatcherKonolige(roi, roi, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1, W, H));
            if (argc >= 13) bm->setROI1(cv::Rect(atoi(argv[9]), atoi(argv[10]), atoi(argv[11]), atoi(argv[12])));
            cv::Mat disp;
            if (bm->compute(lv, rv, disp) != 0) return 4;
            if (bm->compute(lv, rv, disp) != 0) return 4;      // reused output Mat, second frame
            std::vector<short> out((size_t)W * H);
            for (int y = 0; y < H; y++) std::memcpy(&out[(size_t)y * W], disp.ptr<short>(y), (size_t)W * 2);
            dump(argv[8], out.data(), out.size() * 2);
            delete bm;
            return 0;
        }
        if (!std::strcmp(argv[1], "sgbm") && argc >= 10) {
            int W = atoi(argv[2]), H = atoi(argv[3]), nd = atoi(argv[4]), bs = atoi(argv[5]), mode = atoi(argv[9]);
            std::vector<unsigned char> l = slurp(argv[6], (size_t)W * H), r = slurp(argv[7], (size_t)W * H);
            cv::Mat lm(H, W, CV_8UC1, l.data()), rm(H, W, CV_8UC1, r.data()), disp;
            BlockMatcher *bm = new CUDASemiGlobalMatcher(bs, 0, nd, 10, 100, 32, 1, mode, W, H);
            bm->setROI1(cv::Rect(1, 2, 3, 4));
            if (bm->compute(lm, rm, disp) != 0) return 4;
            dump(argv[8], disp.data, (size_t)W * H * 2);
            delete bm;
            return 0;
        }
        if (!std::strcmp(argv[1], "morph") && argc >= 6) {
            int W = atoi(argv[2]), H = atoi(argv[3]);
            std::vector<unsigned char> in = slurp(argv[4], (size_t)W * H);
            VideoFilterDevice *f = new CUDAMorphologicalFilter(W, H, 8);
            if (f->getFrameSize() != W * H || f->getWidth() != W || f->getHeight() != H || f->getBpp() != 8) return 5;
            // Estimator wraps the plugin-owned buffers in Mats and writes the mask into the input buffer
            cv::Mat fin(H, W, CV_8UC1, f->getVideoInBuffer()), fout(H, W, CV_8UC1, f->getVideoOutBuffer());
            std::memcpy(fin.data, in.data(), in.size());
            if (f->run(fin, fout) != 0) return 4;
            dump(argv[5], fout.data, (size_t)W * H);
            delete f;
            return 0;
        }
        if (!std::strcmp(argv[1], "rectify") && argc >= 8) {
            // rectify W H rgb map1 map2 out   (ROI fixed: 2 px margin)
            int W = atoi(argv[2]), H = atoi(argv[3]);
            std::vector<unsigned char> rgb = slurp(argv[4], (size_t)W * H * 3), m1 = slurp(argv[5], (size_t)W * H * 4), m2 = slurp(argv[6], (size_t)W * H * 2);
            cv::Mat img(H, W, CV_8UC1, rgb.data(), (size_t)W * 3);            // the shim has no 3-channel type: only data/step are used
            cv::Mat map1(H, W, CV_16SC1, m1.data(), (size_t)W * 4), map2(H, W, CV_16SC1, m2.data(), (size_t)W * 2);
            CUDARectifier r(map1, map2, cv::Rect(2, 2, W - 4, H - 4));
            cv::Mat rect;
            if (r.run(img, rect) != 0) return 4;
            std::vector<unsigned char> out((size_t)(W - 4) * (H - 4));
            for (int y = 0; y < H - 4; y++) std::memcpy(&out[(size_t)y * (W - 4)], rect.ptr<unsigned char>(y), (size_t)W - 4);
            dump(argv[7], out.data(), out.size());
            return 0;
        }
        if (!std::strcmp(argv[1], "mask") && argc >= 9) {
            // mask W H rgb map1 map2 mask_out boxes_out   (ROI fixed: 2 px margin; HSV range and minObjSize fixed)
            int W = atoi(argv[2]), H = atoi(argv[3]);
            std::vector<unsigned char> rgb = slurp(argv[4], (size_t)W * H * 3), m1 = slurp(argv[5], (size_t)W * H * 4), m2 = slurp(argv[6], (size_t)W * H * 2);
            cv::Mat img(H, W, CV_8UC1, rgb.data(), (size_t)W * 3);
            cv::Mat map1(H, W, CV_16SC1, m1.data(), (size_t)W * 4), map2(H, W, CV_16SC1, m2.data(), (size_t)W * 2);
            const cv::Rect roif(2, 2, W - 4, H - 4);
            CUDAColorMask cm(map1, map2, roif);
            // like Estimator: filter_in / filter_out wrap the filter plugin's buffers (estimator.cpp:141-142)
            VideoFilterDevice *f = new CUDAMorphologicalFilter(roif.width, roif.height, 8);
            cv::Mat fin(roif.height, roif.width, CV_8UC1, f->getVideoInBuffer()), fout(roif.height, roif.width, CV_8UC1, f->getVideoOutBuffer());
            const int low[3] = {20, 40, 40}, high[3] = {130, 255, 255};
            if (cm.run(img, low, high, fin) != 0) return 4;
            dump(argv[7], fin.data, (size_t)roif.width * roif.height);
            if (f->run(fin, fout) != 0) return 4;
            CUDAObjectRegions regions(roif.width, roif.height, 1024);
            std::vector<cv::Rect> bounds; cv::Rect span;
            const int nc = regions.run(fout, 60, bounds, span);
            if (nc < 0) return 4;
            FILE *o = std::fopen(argv[8], "w");
            if (!o) return 2;
            std::fprintf(o, "%d %d %d %d %d\n", nc, span.x, span.y, span.width, span.height);
            for (size_t i = 0; i < bounds.size(); i++) std::fprintf(o, "%d %d %d %d\n", bounds[i].x, bounds[i].y, bounds[i].width, bounds[i].height);
            std::fclose(o);
            delete f;
            return 0;
        }
        if (!std::strcmp(argv[1], "depth") && argc >= 11) {
            int W = atoi(argv[2]), H = atoi(argv[3]);
            std::vector<unsigned char> d = slurp(argv[4], (size_t)W * H * 2), m = slurp(argv[5], (size_t)W * H);
            cv::Mat disp(H, W, CV_16SC1, d.data()), mask(H, W, CV_8UC1, m.data());
            const double Q[16] = {1, 0, 0, -W / 2.0 + 0.37, 0, 1, 0, -H / 2.0 - 0.21, 0, 0, 0, 0.8 * W, 0, 0, 1 / 119.87, 0.004};
            std::vector<cv::Rect> regions;
            regions.push_back(cv::Rect(atoi(argv[7]), atoi(argv[8]), atoi(argv[9]), atoi(argv[10])));
            regions.push_back(cv::Rect(0, 0, W, H));
            CUDADepthEpilogue ep(W, H, 8);
            std::vector<double> mean; std::vector<int> cnt;
            if (ep.run(disp, Q, mask, regions, mean, &cnt) != 0) return 4;
            FILE *f = std::fopen(argv[6], "w");
            if (!f) return 2;
            for (size_t i = 0; i < mean.size(); i++) std::fprintf(f, "%.17g %d\n", mean[i], cnt[i]);
            std::fclose(f);
            return 0;
        }
    } catch (const std::exception &e) {
        std::printf("nodevice: %s\n", e.what());
        return 3;
    }
    return 2;
}
