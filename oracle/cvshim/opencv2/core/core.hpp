// cvshim -- minimal OpenCV-compatible surface (TEST INFRASTRUCTURE ONLY).
//
// Just enough of cv:: for the reference's R21/src/ORBextractor.cc to compile VERBATIM from
// /root/reference (oracle/Makefile, target _ref).  The image primitives forward to the
// integer models in oracle/orb_oracle.cc, which are pinned bit-for-bit against cv2 4.13.0
// (tests/test_oracle_primitives.py).  This is not OpenCV and implements only CV_8UC1 / the
// calls that file makes.
#ifndef CVSHIM_CORE_HPP
#define CVSHIM_CORE_HPP

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <list>
#include <new>
#include <utility>
#include <vector>
#include <chrono>

#include <sys/mman.h>

#include "orb_oracle.h"

// Per-thread time spent inside the shim primitives (seconds): [0] resize, [1] copyMakeBorder, [2] GaussianBlur, [3] FAST.
// bench.py reads them through orbref_stage_times() to time cv2's own primitives beside the shim's.
inline double* cvshim_stage_acc() { static thread_local double acc[4] = {0, 0, 0, 0}; return acc; }
struct CvshimStageTimer {
    int k; std::chrono::steady_clock::time_point t0;
    explicit CvshimStageTimer(int kk) : k(kk), t0(std::chrono::steady_clock::now()) {}
    ~CvshimStageTimer() { cvshim_stage_acc()[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// ---------------------------------------------------------------------------------------------
// Canonical tie-break for DistributeOctTree's sort of (size, ExtractorNode*) pairs
// (R21 ORBextractor.cc:684).  std::list<ExtractorNode> nodes come from a
// per-thread monotonic arena, so "larger pointer" == "created later" regardless of malloc.
// The arena rewinds when the last live list node is freed (end of DistributeOctTree).
// ---------------------------------------------------------------------------------------------
namespace ORB_SLAM2 { class ExtractorNode; }
namespace cvshim_detail {
struct NodeArena {
    char* base; size_t off, cap; long live;
    NodeArena() : base(nullptr), off(0), cap((size_t)1 << 28), live(0) {}
    void* take(size_t bytes) {
        if (!base) {
            void* p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (p == MAP_FAILED) abort();
            base = (char*)p;
        }
        bytes = (bytes + 15) & ~(size_t)15;
        if (off + bytes > cap) abort();
        void* r = base + off;
        off += bytes;
        live++;
        return r;
    }
    void give() { if (--live == 0) off = 0; }
    ~NodeArena() { if (base) munmap(base, cap); }
};
inline NodeArena& node_arena() { static thread_local NodeArena a; return a; }
}  // namespace cvshim_detail

namespace std {
template <> class allocator<_List_node<ORB_SLAM2::ExtractorNode> > {
public:
    typedef _List_node<ORB_SLAM2::ExtractorNode> value_type;
    typedef value_type* pointer;
    typedef const value_type* const_pointer;
    typedef size_t size_type;
    typedef ptrdiff_t difference_type;
    template <class U> struct rebind { typedef allocator<U> other; };
    allocator() noexcept {}
    allocator(const allocator&) noexcept {}
    template <class U> allocator(const allocator<U>&) noexcept {}
    // template so that sizeof(value_type) is only needed at the call site (ExtractorNode complete there)
    template <class V = value_type> pointer allocate(size_type n, const void* = nullptr) {
        return (pointer)cvshim_detail::node_arena().take(n * sizeof(V));
    }
    void deallocate(pointer, size_type) { cvshim_detail::node_arena().give(); }
    template <class U, class... Args> void construct(U* p, Args&&... args) { ::new ((void*)p) U(std::forward<Args>(args)...); }
    template <class U> void destroy(U* p) { p->~U(); }
    size_type max_size() const noexcept { return (size_t)-1 / 256; }
    bool operator==(const allocator&) const { return true; }
    bool operator!=(const allocator&) const { return false; }
};
}  // namespace std

namespace cv {

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> 3) & 63) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << 3))
#define CV_PI 3.1415926535897932384626433832795

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)floor(v); }
inline int cvFloor(float v) { return (int)floorf(v); }
inline int cvCeil(double v) { return (int)ceil(v); }
inline int cvCeil(float v) { return (int)ceilf(v); }
inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> inline Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
    size_t p;
    MatStep() : p(0) {}
    operator size_t() const { return p; }
    MatStep& operator=(size_t s) { p = s; return *this; }
};

class Mat;
// what cv::Mat::zeros / ones / eye return (a cv::MatExpr in OpenCV): assigned into an existing header it fills in place
struct MatConst {
    int rows, cols, type, kind;      // kind 0 zeros, 1 ones, 2 eye
    operator Mat() const;
};
typedef MatConst MatZeros;

// Reference-counted matrix header with ROI views.  CV_8UC1 for the image path (the extractor); CV_32F / CV_64F (1 or 2
// channels) with the handful of arithmetic operators the reference's Frame / KeyFrame / MapPoint / ORBmatcher use.
class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;

    Mat() : rows(0), cols(0), data(nullptr), type_(CV_8UC1), buf_(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), type_(CV_8UC1), buf_(nullptr) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(nullptr), type_(CV_8UC1), buf_(nullptr) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), data((uchar*)ext), type_(type), buf_(nullptr) {
        step = st ? st : (size_t)c * elemSize();
    }
    Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), type_(m.type_), buf_(m.buf_) { if (buf_) buf_->ref++; }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) {
        if (this != &m) {
            if (m.buf_) m.buf_->ref++;
            release();
            rows = m.rows; cols = m.cols; data = m.data; step = m.step; type_ = m.type_; buf_ = m.buf_;
        }
        return *this;
    }
    // cv::Mat::operator=(const MatExpr&) for Mat::zeros / ones / eye: create() is a no-op on a matching header, then fill.
    Mat& operator=(const MatConst& z) {
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; r++) memset(data + (size_t)r * step.p, 0, (size_t)cols * elemSize());
        if (z.kind) {
            for (int r = 0; r < rows; r++)
                for (int c = 0; c < cols; c++)
                    if (z.kind == 1 || r == c) set_scalar(r, c, 1.0);
        }
        return *this;
    }
    static MatConst zeros(int r, int c, int type) { MatConst z = {r, c, type, 0}; return z; }
    static MatConst ones(int r, int c, int type) { MatConst z = {r, c, type, 1}; return z; }
    static MatConst eye(int r, int c, int type) { MatConst z = {r, c, type, 2}; return z; }

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && type_ == type) return;
        release();
        type_ = type;
        rows = r; cols = c; step = (size_t)c * elemSize();
        if ((size_t)r * c > 0) {
            buf_ = new Buf;
            buf_->ref = 1;
            buf_->mem = (uchar*)malloc((size_t)r * c * elemSize());
            data = buf_->mem;
        }
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() {
        if (buf_ && --buf_->ref == 0) { free(buf_->mem); delete buf_; }
        buf_ = nullptr; data = nullptr; rows = cols = 0; step = 0;
    }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize1() const { const int d = depth(); return d == CV_8U ? 1 : (d == CV_64F ? 8 : 4); }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step.p / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step.p == (size_t)cols * elemSize() || rows <= 1; }

    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step.p + (size_t)r.x * elemSize();
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step.p, data + (size_t)r * step.p, (size_t)cols * elemSize());
        return m;
    }
    // cv::Mat::copyTo: dst.create() keeps a matching header (so a ROI view of another matrix is written in place)
    void copyTo(Mat& dst) const {
        dst.create(rows, cols, type_);
        for (int r = 0; r < rows; r++) memmove(dst.data + (size_t)r * dst.step.p, data + (size_t)r * step.p, (size_t)cols * elemSize());
    }
    void copyTo(const Mat& dst_view) const { Mat d(dst_view); copyTo(d); }      // a temporary ROI: Rwc.copyTo(Twc.rowRange(...))
    // same memory, other channel count (Frame::UndistortKeyPoints: N x 2 one-channel <-> N x 1 two-channel)
    Mat reshape(int cn) const {
        Mat m(*this);
        const int w = cols * channels();
        assert(isContinuous() || rows == 1);
        assert(w % cn == 0);
        m.type_ = CV_MAKETYPE(depth(), cn);
        m.cols = w / cn;
        return m;
    }
    void convertTo(Mat& dst, int rtype) const {
        assert(channels() == 1);
        Mat out(rows, cols, CV_MAT_DEPTH(rtype));
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) out.set_scalar(r, c, get_scalar(r, c));
        dst = out;
    }
    Mat t() const {
        Mat m(cols, rows, type_);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) memcpy(m.data + (size_t)c * m.step.p + (size_t)r * elemSize(), data + (size_t)r * step.p + (size_t)c * elemSize(), elemSize());
        return m;
    }
    double dot(const Mat& b) const {
        assert(total() == b.total());
        double acc = 0;
        const int n = (int)total();
        for (int i = 0; i < n; i++) acc += lin_scalar(i) * b.lin_scalar(i);
        return acc;
    }
    uchar* ptr(int r = 0) { return data + (ptrdiff_t)r * (ptrdiff_t)step.p; }
    const uchar* ptr(int r = 0) const { return data + (ptrdiff_t)r * (ptrdiff_t)step.p; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p + (ptrdiff_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const {
        return *(const T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p + (ptrdiff_t)c * sizeof(T));
    }
    // cv::Mat::at(int i0): element i0 of a row / column vector (or of a continuous matrix)
    template <typename T> T& at(int i) { return const_cast<T&>(static_cast<const Mat*>(this)->at<T>(i)); }
    template <typename T> const T& at(int i) const {
        if (rows == 1 || isContinuous() && cols != 1) return *(const T*)(data + (ptrdiff_t)i * sizeof(T));
        if (cols == 1) return *(const T*)(data + (ptrdiff_t)i * (ptrdiff_t)step.p);
        const int r = i / cols;
        return at<T>(r, i - r * cols);
    }
    // element access as double, whatever the depth (one channel)
    double get_scalar(int r, int c) const {
        const uchar* p = data + (size_t)r * step.p + (size_t)c * elemSize();
        switch (depth()) { case CV_8U: return *p; case CV_32F: return *(const float*)p; case CV_64F: return *(const double*)p; default: return *(const int*)p; }
    }
    void set_scalar(int r, int c, double v) {
        uchar* p = data + (size_t)r * step.p + (size_t)c * elemSize();
        switch (depth()) { case CV_8U: *p = (uchar)v; break; case CV_32F: *(float*)p = (float)v; break; case CV_64F: *(double*)p = v; break; default: *(int*)p = (int)v; }
    }
    double lin_scalar(int i) const { const int w = cols * channels(); Mat one = channels() == 1 ? *this : reshape(1); return one.get_scalar(i / w, i % w); }

private:
    struct Buf { int ref; uchar* mem; };
    int type_;
    Buf* buf_;
};
inline MatConst::operator Mat() const { Mat m; m = *this; return m; }

// ---- the arithmetic the reference writes on pose / point matrices (CV_32F).  Products and sums are formed in double
// and rounded once per element, like cv::gemm's float path (double accumulators).
inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows && a.type() == b.type());
    Mat m(a.rows, b.cols, a.type());
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            double acc = 0;
            for (int k = 0; k < a.cols; k++) acc += a.get_scalar(r, k) * b.get_scalar(k, c);
            m.set_scalar(r, c, acc);
        }
    return m;
}
inline Mat cvshim_binary(const Mat& a, const Mat& b, double sb) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat m(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) m.set_scalar(r, c, a.get_scalar(r, c) + sb * b.get_scalar(r, c));
    return m;
}
inline Mat cvshim_scaled(const Mat& a, double s) {
    Mat m(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) m.set_scalar(r, c, a.get_scalar(r, c) * s);
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return cvshim_binary(a, b, 1.0); }
inline Mat operator-(const Mat& a, const Mat& b) { return cvshim_binary(a, b, -1.0); }
inline Mat operator-(const Mat& a) { return cvshim_scaled(a, -1.0); }
inline Mat operator*(const Mat& a, double s) { return cvshim_scaled(a, s); }
inline Mat operator*(double s, const Mat& a) { return cvshim_scaled(a, s); }
inline Mat operator/(const Mat& a, double s) { return cvshim_scaled(a, 1.0 / s); }       // MatExpr: alpha = 1/s

inline double norm(const Mat& a, int type = NORM_L2) {
    double acc = 0;
    const int n = (int)a.total() * a.channels();
    for (int i = 0; i < n; i++) { const double v = a.lin_scalar(i); acc += type == NORM_L1 ? std::fabs(v) : v * v; }
    return type == NORM_L1 ? acc : std::sqrt(acc);
}
inline double norm(const Mat& a, const Mat& b, int type = NORM_L2) {
    assert(a.rows == b.rows && a.cols == b.cols);
    double acc = 0;
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols; c++) { const double v = a.get_scalar(r, c) - b.get_scalar(r, c); acc += type == NORM_L1 ? std::fabs(v) : v * v; }
    return type == NORM_L1 ? acc : std::sqrt(acc);
}

// cv::Mat_<T>(r, c) << a, b, c ...   (Frame::UnprojectStereo, KeyFrame::SetPose)
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<T>* m, T v);
    MatCommaInitializer_& operator,(T v);
    operator Mat() const;
    operator Mat_<T>() const;
private:
    Mat_<T>* m_; int i_;
};
template <typename T> struct cvshim_depth;
template <> struct cvshim_depth<float> { enum { value = CV_32F }; };
template <> struct cvshim_depth<double> { enum { value = CV_64F }; };
template <> struct cvshim_depth<uchar> { enum { value = CV_8U }; };
template <> struct cvshim_depth<int> { enum { value = CV_32S }; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, cvshim_depth<T>::value) {}
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
};
template <typename T> MatCommaInitializer_<T>::MatCommaInitializer_(Mat_<T>* m, T v) : m_(m), i_(0) { (*this), v; }
template <typename T> MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(T v) {
    m_->template at<T>(i_ / m_->cols, i_ % m_->cols) = v; i_++;
    return *this;
}
template <typename T> MatCommaInitializer_<T>::operator Mat() const { return *m_; }
template <typename T> MatCommaInitializer_<T>::operator Mat_<T>() const { return *m_; }
template <typename T, typename V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
    return MatCommaInitializer_<T>(const_cast<Mat_<T>*>(&m), (T)v);
}

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    const Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() : o_(nullptr) {}
    _OutputArray(Mat& m) : _InputArray(m), o_(&m) {}
    _OutputArray(const Mat& m) : _InputArray(m), o_(const_cast<Mat*>(&m)) {}
    void create(int r, int c, int type) const { o_->create(r, c, type); }
    void create(Size s, int type) const { o_->create(s, type); }
    void release() const { if (o_) o_->release(); }
    Mat getMat() const { return *o_; }
    Mat& getMatRef() const { return *o_; }
private:
    Mat* o_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline InputArray noArray() { static _InputArray a; return a; }

}  // namespace cv

#endif
