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
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

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

struct MatZeros { int rows, cols, type; };

// CV_8UC1-only reference-counted matrix header with ROI views.
class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;

    Mat() : rows(0), cols(0), data(nullptr), buf_(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), buf_(nullptr) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(nullptr), buf_(nullptr) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), data((uchar*)ext), buf_(nullptr) {
        (void)type; step = st ? st : (size_t)c;
    }
    Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), buf_(m.buf_) { if (buf_) buf_->ref++; }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) {
        if (this != &m) {
            if (m.buf_) m.buf_->ref++;
            release();
            rows = m.rows; cols = m.cols; data = m.data; step = m.step; buf_ = m.buf_;
        }
        return *this;
    }
    // cv::Mat::operator=(const MatExpr&) for Mat::zeros: create() is a no-op on a matching header, then fill.
    Mat& operator=(const MatZeros& z) {
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; r++) memset(data + (size_t)r * step.p, 0, cols);
        return *this;
    }
    static MatZeros zeros(int r, int c, int type) { MatZeros z = {r, c, type}; return z; }

    void create(int r, int c, int type) {
        assert(type == CV_8UC1);
        (void)type;
        if (data && rows == r && cols == c) return;
        release();
        rows = r; cols = c; step = (size_t)c;
        if ((size_t)r * c > 0) {
            buf_ = new Buf;
            buf_->ref = 1;
            buf_->mem = (uchar*)malloc((size_t)r * c);
            data = buf_->mem;
        }
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() {
        if (buf_ && --buf_->ref == 0) { free(buf_->mem); delete buf_; }
        buf_ = nullptr; data = nullptr; rows = cols = 0; step = 0;
    }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    int type() const { return CV_8UC1; }
    int channels() const { return 1; }
    size_t step1() const { return step.p; }
    size_t elemSize() const { return 1; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step.p == (size_t)cols || rows <= 1; }

    Mat operator()(const Rect& r) const {
        Mat m(*this);
        m.data = data + (size_t)r.y * step.p + r.x;
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step.p, data + (size_t)r * step.p, cols);
        return m;
    }
    uchar* ptr(int r = 0) { return data + (ptrdiff_t)r * (ptrdiff_t)step.p; }
    const uchar* ptr(int r = 0) const { return data + (ptrdiff_t)r * (ptrdiff_t)step.p; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p); }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p + (ptrdiff_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const {
        return *(const T*)(data + (ptrdiff_t)r * (ptrdiff_t)step.p + (ptrdiff_t)c * sizeof(T));
    }

private:
    struct Buf { int ref; uchar* mem; };
    Buf* buf_;
};

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
