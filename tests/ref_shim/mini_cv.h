// mini_cv.h -- TEST INFRASTRUCTURE. Stand-in for the slice of the OpenCV C++ API that the reference's hot-path translation
// units touch, so that they compile UNMODIFIED into oracle/_ref/libdsdtm_ref.so (OpenCV's C++ library is not installed).
//
//   * cv::Mat is a plain ref-counted 2-D byte buffer with the members the reference reads (data, rows, cols, step.p[0],
//     at<>, ptr<>); Point2f -> Point conversion rounds like OpenCV's saturate_cast (cvRound);
//   * the imgproc calls on the path (pyrDown, circle, undistortPoints) are served by the oracle's restatements, which are
//     pinned against cv2 4.13 golden vectors (tests/golden/*.npz); threshold / saturating subtraction are trivial;
//   * drawing / GUI calls are no-ops; cv::FileStorage reads flat "key: value" YAML.
// Nothing under dsdtm_b200/ includes this file.
#ifndef MINI_CV_H
#define MINI_CV_H

#include <emmintrin.h>

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../oracle/dsdtm_oracle.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_GRAY2BGR 8
#define CV_THRESH_BINARY 0

// OpenCV on x86-64: _mm_cvtsd_si32 (round to nearest even under the default MXCSR; NaN / overflow -> INT_MIN)
inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }

namespace cv {

template <class T> inline T saturate_cast(float v) { return T(v); }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <class T> inline T saturate_cast(double v) { return T(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <class T> inline T saturate_cast(int v) { return T(v); }

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <class U> operator Point_<U>() const { return Point_<U>(saturate_cast<U>(x), saturate_cast<U>(y)); }
};
template <class T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> inline Point_<T> operator*(int s, const Point_<T>& a) { return Point_<T>(T(s * a.x), T(s * a.y)); }
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double operator[](int i) const { return val[i]; }
};

struct KeyPoint {
    Point2f pt;
    float size;
    KeyPoint() : size(0) {}
    KeyPoint(float x, float y, float s) : pt(x, y), size(s) {}
};

struct MatStep {
    size_t buf[2];
    size_t* p;
    MatStep() : p(buf) { buf[0] = buf[1] = 0; }
    MatStep(const MatStep& o) : p(buf) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; }
    MatStep& operator=(const MatStep& o) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; return *this; }
    operator size_t() const { return buf[0]; }
};

class Mat {
public:
    int flags, dims, rows, cols;
    uchar* data;
    MatStep step;

    Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr) {}
    Mat(int r, int c, int type) : flags(0), dims(2), rows(0), cols(0), data(nullptr) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) : flags(0), dims(2), rows(0), cols(0), data(nullptr) { create(r, c, type); fill(s); }
    Mat(Size sz, int type, const Scalar& s) : flags(0), dims(2), rows(0), cols(0), data(nullptr) { create(sz.height, sz.width, type); fill(s); }
    // a header over memory the caller owns
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : flags(type), dims(2), rows(r), cols(c), data(static_cast<uchar*>(ext))
    {
        step.buf[0] = stp ? stp : (size_t)c * elemSize();
        step.buf[1] = elemSize();
    }

    void create(int r, int c, int type)
    {
        flags = type; rows = r; cols = c;
        step.buf[1] = elemSize();
        step.buf[0] = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step.buf[0] + 64, (uchar)0);  // slack: the reference reads 1 past the edge (Q4)
        data = buf_->data();
    }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> 3) + 1; }
    size_t elemSize1() const { static const int s[7] = { 1, 1, 2, 2, 4, 4, 8 }; return (size_t)s[depth()]; }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    bool isContinuous() const { return step.buf[0] == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; }

    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step.buf[0] + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step.buf[0] + (size_t)c * sizeof(T)); }
    template <class T> T& at(int i) { return (rows == 1) ? at<T>(0, i) : at<T>(i, 0); }
    template <class T> const T& at(int i) const { return (rows == 1) ? at<T>(0, i) : at<T>(i, 0); }
    template <class T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <class T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step.buf[0]); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step.buf[0]); }

    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step.buf[0], data + (size_t)r * step.buf[0], (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    Mat reshape(int cn) const  // continuous data only: same bytes, new channel count
    {
        Mat m(*this);
        const size_t row_bytes = (size_t)cols * elemSize();
        m.flags = CV_MAKETYPE(depth(), cn);
        m.cols = (int)(row_bytes / m.elemSize());
        m.step.buf[1] = m.elemSize();
        return m;
    }
    // CV_16U / CV_32F -> CV_32F with a scale, the only form the reference uses (ref: src/Tracking.cpp:56): one float multiply per pixel
    // (a single correctly rounded product: nothing to restate)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0) const
    {
        if (rtype != CV_32F || channels() != 1 || (depth() != CV_16U && depth() != CV_32F)) std::abort();
        Mat out(rows, cols, CV_32FC1);
        const float a = (float)alpha;
        for (int r = 0; r < rows; ++r) {
            for (int c = 0; c < cols; ++c) out.ptr<float>(r)[c] = (depth() == CV_16U ? (float)ptr<ushort>(r)[c] : ptr<float>(r)[c]) * a;
        }
        dst = out;
    }
    static Mat eye(int r, int c, int type)
    {
        Mat m(r, c, type, Scalar(0));
        for (int i = 0; i < (r < c ? r : c); ++i) {
            if (m.depth() == CV_32F) m.at<float>(i, i) = 1.f;
            else if (m.depth() == CV_64F) m.at<double>(i, i) = 1.0;
            else m.at<uchar>(i, i) = 1;
        }
        return m;
    }

private:
    void fill(const Scalar& s)
    {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * channels(); ++c) {
                const double v = s.val[c % channels()];
                switch (depth()) {
                case CV_8U: ptr<uchar>(r)[c] = (uchar)v; break;
                case CV_16U: ptr<ushort>(r)[c] = (ushort)v; break;
                case CV_32S: ptr<int>(r)[c] = (int)v; break;
                case CV_32F: ptr<float>(r)[c] = (float)v; break;
                case CV_64F: ptr<double>(r)[c] = v; break;
                default: std::abort();
                }
            }
    }
    std::shared_ptr<std::vector<uchar> > buf_;
};

// CV_8UC1 only: saturating subtraction (ref: src/Frame.cpp:296)
inline Mat operator-(const Mat& a, const Mat& b)
{
    Mat r(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) {
            const int v = (int)a.at<uchar>(y, x) - (int)b.at<uchar>(y, x);
            r.at<uchar>(y, x) = (uchar)(v < 0 ? 0 : v);
        }
    return r;
}

inline Mat noArray() { return Mat(); }

// cv::pyrDown on CV_8UC1, default size / border: the oracle's restatement, pinned against cv2 4.13 (tests/golden/pyrdown_cv2.npz)
inline void pyrDown(const Mat& src, Mat& dst)
{
    Mat out((src.rows + 1) / 2, (src.cols + 1) / 2, CV_8UC1);
    orc_pyrdown_u8(src.data, src.cols, src.rows, (int)src.step.buf[0], out.data);
    dst = out;
}
// cv::circle(img, center, radius, color, -1): the oracle's restatement, pinned against cv2 4.13 (tests/golden/circle_cv2.npz)
inline void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1)
{
    if (thickness >= 0) std::abort();  // the reference only draws filled discs on the path
    orc_circle_fill(img.data, img.cols, img.rows, (int)img.step.buf[0], center.x, center.y, radius, (uchar)color.val[0]);
}
inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type)
{
    if (type != CV_THRESH_BINARY) std::abort();
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) out.at<uchar>(y, x) = src.at<uchar>(y, x) > thresh ? (uchar)maxval : (uchar)0;
    dst = out;
    return thresh;
}
// cv::undistortPoints(src, dst, K, dist, noArray(), P) with CV_32FC2 points and CV_32F K / dist / P == K, the only form the reference
// calls (ref: src/Frame.cpp:121-122): the oracle's restatement, pinned against cv2 4.13 (tests/golden/undistort_cv2.npz)
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat&, const Mat& P)
{
    orc_cam cam;
    cam.width = cam.height = 0;
    cam.fx = K.at<float>(0, 0); cam.fy = K.at<float>(1, 1); cam.cx = K.at<float>(0, 2); cam.cy = K.at<float>(1, 2); cam.f = cam.fx;
    if (P.at<float>(0, 0) != cam.fx || P.at<float>(1, 1) != cam.fy || P.at<float>(0, 2) != cam.cx || P.at<float>(1, 2) != cam.cy) std::abort();
    float d[5];
    for (int i = 0; i < 5; ++i) d[i] = dist.at<float>(i);
    const int n = src.rows * src.cols;
    std::vector<float> out((size_t)2 * n + 2);
    orc_undistort_points(&cam, d, src.ptr<float>(0), n, out.data());
    Mat o(src.rows, src.cols, CV_32FC2);
    std::memcpy(o.data, out.data(), (size_t)2 * n * sizeof(float));
    dst = o;
}
template <class... A> inline void undistort(A&&...) { std::abort(); }  // only in a comment of the reference
template <class... A> inline void cvtColor(A&&...) {}
template <class... A> inline void rectangle(A&&...) {}
template <class... A> inline void line(A&&...) {}
template <class... A> inline void namedWindow(A&&...) {}
template <class... A> inline void imshow(A&&...) {}
inline int waitKey(int = 0) { return -1; }

// flat "key: value" YAML (the layout of the reference's Config/*.yaml files; '%YAML', '---', comments and git conflict
// markers are skipped)
class FileNode {
public:
    FileNode() : ok_(false) {}
    explicit FileNode(const std::string& v) : ok_(true), v_(v) {}
    operator int() const { return ok_ ? (int)std::strtod(v_.c_str(), nullptr) : 0; }
    operator float() const { return ok_ ? (float)std::strtod(v_.c_str(), nullptr) : 0.f; }
    operator double() const { return ok_ ? std::strtod(v_.c_str(), nullptr) : 0.0; }
    operator std::string() const { return v_; }
    bool empty() const { return !ok_; }
private:
    bool ok_;
    std::string v_;
};

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() : open_(false) {}
    FileStorage(const std::string& file, int) : open_(false)
    {
        std::ifstream in(file.c_str());
        if (!in) return;
        open_ = true;
        std::string ln;
        while (std::getline(in, ln)) {
            const size_t h = ln.find('#');
            if (h != std::string::npos) ln.erase(h);
            const size_t c = ln.find(':');
            if (c == std::string::npos || ln.empty() || ln[0] == '%' || ln[0] == '-' || ln[0] == '<' || ln[0] == '=' || ln[0] == '>') continue;
            std::string k = trim(ln.substr(0, c)), v = trim(ln.substr(c + 1));
            if (v.size() >= 2 && v[0] == '"' && v[v.size() - 1] == '"') v = v.substr(1, v.size() - 2);
            if (!k.empty()) kv_[k] = v;
        }
    }
    bool isOpened() const { return open_; }
    void release() { open_ = false; kv_.clear(); }
    FileNode operator[](const std::string& k) const
    {
        std::map<std::string, std::string>::const_iterator it = kv_.find(k);
        return it == kv_.end() ? FileNode() : FileNode(it->second);
    }
private:
    static std::string trim(const std::string& s)
    {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    bool open_;
    std::map<std::string, std::string> kv_;
};

}  // namespace cv

#endif
