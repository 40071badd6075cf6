// cv_compat.h -- the few OpenCV core types the plugin interfaces use.
//
// The reference's plugin ABCs (include/stereo-matcher/stereo-matcher.h:13-19, include/filter/filter.h:13-37)
// are written against cv::InputArray / cv::OutputArray / cv::Rect.  Where OpenCV's headers exist this file
// simply includes them.  Where they do not (this build image has no OpenCV C++ headers) it provides a
// minimal stand-in with the same spelling so that the adapters compile and can be unit-tested; the
// stand-in is NOT an OpenCV replacement.
#pragma once

#if defined(RTDM_USE_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core.hpp>))
#include <opencv2/core.hpp>
#define RTDM_HAVE_OPENCV 1
#else
#define RTDM_HAVE_OPENCV 0
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define CV_8UC1 0
#define CV_16SC1 3

namespace cv {

struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};

// Minimal single-channel matrix: owns its buffer unless constructed around user data.
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char *data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *user, size_t step_ = 0)
        : rows(r), cols(c), step(step_ ? step_ : (size_t)c * elem(type)), data((unsigned char *)user), type_(type) {}
    Mat(const Mat &o) { *this = o; }
    Mat &operator=(const Mat &o)
    {
        if (this == &o) return *this;
        release();
        rows = o.rows; cols = o.cols; step = o.step; type_ = o.type_;
        if (o.owned_) {
            owned_ = (unsigned char *)std::malloc(step * rows);
            std::memcpy(owned_, o.owned_, step * rows);
            data = owned_;
        } else data = o.data;
        return *this;
    }
    ~Mat() { release(); }
    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;   // same size: no realloc (like OpenCV)
        release();
        rows = r; cols = c; type_ = type; step = (size_t)c * elem(type);
        owned_ = (unsigned char *)std::malloc(step * (size_t)r);
        data = owned_;
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + step * (size_t)r); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + step * (size_t)r); }
    // ROI view (non-owning), like cv::Mat::operator()(Rect)
    Mat operator()(const Rect &r) const { return Mat(r.height, r.width, type_, data + step * (size_t)r.y + (size_t)r.x * elem(type_), step); }

private:
    static size_t elem(int type) { return type == CV_16SC1 ? 2 : 1; }
    void release() { if (owned_) std::free(owned_); owned_ = nullptr; data = nullptr; }
    unsigned char *owned_ = nullptr;
    int type_ = CV_8UC1;
};

// In real OpenCV these are proxy classes; the plugins only ever pass cv::Mat through them.
struct _InputArray {
    const Mat *m;
    _InputArray(const Mat &mat) : m(&mat) {}
    const Mat &getMat() const { return *m; }
};
struct _OutputArray {
    Mat *m;
    _OutputArray(Mat &mat) : m(&mat) {}
    void create(int rows, int cols, int type) const { m->create(rows, cols, type); }
    Mat &getMatRef() const { return *m; }
    Mat getMat() const { return Mat(m->rows, m->cols, m->type(), m->data, m->step); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

}  // namespace cv
#endif
