// cvshim — TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// A minimal OpenCV-*shaped* facade, written from scratch, that lets the reference's own hot-path sources
// (/root/reference/src/processing/**, compiled where they lie by oracle/build_ref.py) build in an image that
// has no OpenCV C++ headers.  It declares just the slice of the cv:: API those files use.  Every pixel
// operation is forwarded to the *real* OpenCV kernels through the `cv2` Python module (cvshim.cpp), so the
// arithmetic is OpenCV's, the control flow is the reference's, and nothing here is a re-implementation of an
// image-processing routine.  The exceptions, implemented natively because cv2 does not expose them:
//   * cv::Mat bookkeeping (headers, ref-counted buffers, views, clone/copyTo/reshape/setTo/zeros/t),
//   * cv::Mat::convertTo for the four depth pairs the path uses (documented at its definition),
//   * cv::MatExpr lowering rules (which OpenCV call an expression such as `a*x + b*y` turns into),
//     restated from OpenCV's modules/core/src/matop.cpp behaviour.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr)                                                              \
    do {                                                                             \
        if (!(expr)) throw std::runtime_error("cvshim: CV_Assert failed: " #expr);   \
    } while (0)

namespace cv {

typedef unsigned char uchar;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
template <typename T> bool operator==(const Size_<T>& a, const Size_<T>& b) { return a.width == b.width && a.height == b.height; }
template <typename T> bool operator!=(const Size_<T>& a, const Size_<T>& b) { return !(a == b); }
typedef Size_<int> Size;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() : val{0, 0, 0, 0} {}
    Scalar_(T v0) : val{v0, 0, 0, 0} {}
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) : val{v0, v1, v2, v3} {}
    static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    bool isReal() const { return val[1] == 0 && val[2] == 0 && val[3] == 0; }
};
typedef Scalar_<double> Scalar;

enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum DftFlags { DFT_INVERSE = 1, DFT_SCALE = 2, DFT_ROWS = 4 };
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_MINMAX = 32 };

class MatExpr;

class Mat {
public:
    int flags = 0;   // the type (depth + channels)
    int dims = 2;
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;   // bytes per row

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(const Mat&) = default;                 // header copy, shares the buffer (as in OpenCV)
    Mat(const Mat& m, const Rect& roi);        // view
    Mat& operator=(const Mat&) = default;
    Mat& operator=(const Scalar& s) { setTo(s); return *this; }
    Mat& operator=(const MatExpr& e);

    int type() const { return flags & 0xFFF; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const;
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    size_t total() const { return (size_t)rows * (size_t)cols; }
    Size size() const { return Size(cols, rows); }

    // OpenCV's create(): keeps the buffer when size and type already match, else allocates a fresh one
    void create(int r, int c, int type);
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { *this = Mat(); }

    template <typename T> T* ptr(int row = 0) { return reinterpret_cast<T*>(data + (size_t)row * step); }
    template <typename T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(data + (size_t)row * step); }
    template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }

    Mat clone() const;
    void copyTo(Mat& dst) const;
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s);
    Mat reshape(int cn, int new_rows = 0) const;
    Mat colRange(int startcol, int endcol) const { return Mat(*this, Rect(startcol, 0, endcol - startcol, rows)); }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    MatExpr mul(const Mat& m, double scale = 1) const;
    Mat t() const;

    // cvshim-internal: take over an externally owned buffer (a numpy array returned by cv2)
    void adopt(int r, int c, int type, uchar* ptr, std::shared_ptr<uchar[]> owner);

    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }   // create() zero-fills
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }

private:
    std::shared_ptr<uchar[]> buf_;
};

template <typename T> struct DepthOf;
template <> struct DepthOf<float> { static constexpr int value = CV_32F; };
template <> struct DepthOf<double> { static constexpr int value = CV_64F; };
template <> struct DepthOf<uchar> { static constexpr int value = CV_8U; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() = default;
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(DepthOf<T>::value, 1)) {}
};

template <typename T> class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(const Mat_<T>& m) : m_(m), i_(0) {}
    template <typename T2> MatCommaInitializer_& operator,(T2 v) {
        CV_Assert(i_ < m_.total());
        m_.template ptr<T>((int)(i_ / (size_t)m_.cols))[i_ % (size_t)m_.cols] = (T)v;
        ++i_;
        return *this;
    }
    operator Mat_<T>() const { return m_; }
private:
    Mat_<T> m_;
    size_t i_;
};
template <typename T, typename T2> MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v) {
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}

// Lazy matrix expression with OpenCV's lowering rules (see cvshim.cpp).
class MatExpr {
public:
    enum Kind { ADDEX, MUL };   // ADDEX: a*alpha + b*beta + s ; MUL: a.mul(b) * alpha
    Kind kind = ADDEX;
    Mat a, b;
    double alpha = 1, beta = 0;
    Scalar s;
    MatExpr() = default;
    explicit MatExpr(const Mat& m) : a(m) {}
    operator Mat() const;
    Size size() const { return a.size(); }
    int type() const { return a.type(); }
};

MatExpr operator+(const Mat& a, const Mat& b);
MatExpr operator+(const MatExpr& e, const Mat& m);
MatExpr operator+(const Mat& m, const MatExpr& e);
MatExpr operator+(const MatExpr& e1, const MatExpr& e2);
MatExpr operator-(const Mat& a, const Mat& b);
MatExpr operator-(const MatExpr& e, const Mat& m);
MatExpr operator-(const Mat& m, const MatExpr& e);
MatExpr operator-(const MatExpr& e1, const MatExpr& e2);
MatExpr operator*(const Mat& a, double s);
MatExpr operator*(double s, const Mat& a);
MatExpr operator*(const MatExpr& e, double s);
MatExpr operator*(double s, const MatExpr& e);
MatExpr operator/(const Mat& a, double s);
MatExpr operator/(const MatExpr& e, double s);

// ---- core functions (forwarded to cv2) -----------------------------------------------------------
void add(const Mat& a, const Mat& b, Mat& dst);
void subtract(const Mat& a, const Mat& b, Mat& dst);
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1);
void multiply(const Mat& a, double s, Mat& dst);
void divide(const Mat& a, const Mat& b, Mat& dst, double scale = 1);
void divide(const Mat& a, double s, Mat& dst);
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst);
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst);
void sqrt(const Mat& src, Mat& dst);
void split(const Mat& src, Mat* mv);
void split(const Mat& src, std::vector<Mat>& mv);
void merge(const Mat* mv, size_t count, Mat& dst);
void merge(const std::vector<Mat>& mv, Mat& dst);
void hconcat(const Mat& a, const Mat& b, Mat& dst);
void transpose(const Mat& src, Mat& dst);
void patchNaNs(Mat& a, double val = 0);
void normalize(const Mat& src, Mat& dst, double alpha = 1, double beta = 0, int norm_type = NORM_L2, int dtype = -1);
void minMaxLoc(const Mat& src, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr);
void dft(const Mat& src, Mat& dst, int flags = 0, int nonzeroRows = 0);
void idft(const Mat& src, Mat& dst, int flags = 0, int nonzeroRows = 0);
void mulSpectrums(const Mat& a, const Mat& b, Mat& c, int flags, bool conjB = false);
int getOptimalDFTSize(int vecsize);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());

}  // namespace cv
