// cvshim — TEST INFRASTRUCTURE ONLY (oracle/).  Implementation of the OpenCV-shaped facade in opencv2/*.hpp:
// Mat bookkeeping natively, MatExpr lowering as OpenCV's matop.cpp does it, and every pixel operation forwarded
// to the real OpenCV kernel of the same name through the `cv2` Python module (pybind11; the GIL is held because
// this code only ever runs inside calls made from Python into the _livim_ref extension module).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <climits>
#include <string>

#include "opencv2/core.hpp"
#include "opencv2/imgproc.hpp"

namespace py = pybind11;

namespace cv {

// ---------------------------------------------------------------------------------------------------
// Mat bookkeeping
// ---------------------------------------------------------------------------------------------------
static size_t depth_size(int depth) {
    switch (depth) {
    case CV_8U: case CV_8S: return 1;
    case CV_16U: case CV_16S: return 2;
    case CV_32S: case CV_32F: return 4;
    case CV_64F: return 8;
    }
    throw std::runtime_error("cvshim: unsupported depth");
}

size_t Mat::elemSize1() const { return depth_size(depth()); }

void Mat::create(int r, int c, int type) {
    type &= 0xFFF;
    if (data && rows == r && cols == c && this->type() == type) return;   // OpenCV: reuse matching buffer
    flags = type;
    rows = r;
    cols = c;
    step = (size_t)c * depth_size(CV_MAT_DEPTH(type)) * (size_t)CV_MAT_CN(type);
    const size_t n = step * (size_t)r;
    buf_ = std::shared_ptr<uchar[]>(new uchar[n ? n : 1]());   // zero-filled
    data = buf_.get();
}

Mat::Mat(const Mat& m, const Rect& roi) : Mat(m) {
    CV_Assert(roi.x >= 0 && roi.y >= 0 && roi.width >= 0 && roi.height >= 0 && roi.x + roi.width <= m.cols &&
              roi.y + roi.height <= m.rows);
    data = m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize();
    rows = roi.height;
    cols = roi.width;
}

void Mat::copyTo(Mat& dst) const {
    if (empty()) { dst.release(); return; }
    // a view of dst's own buffer keeps that buffer alive through buf_ while dst re-allocates
    Mat src = *this;
    dst.create(rows, cols, type());
    if (dst.data == src.data && dst.step == src.step) return;
    const size_t rowbytes = (size_t)cols * elemSize();
    for (int y = 0; y < rows; ++y) std::memmove(dst.data + (size_t)y * dst.step, src.data + (size_t)y * src.step, rowbytes);
}

Mat Mat::clone() const {
    Mat m;
    copyTo(m);
    return m;
}

Mat& Mat::setTo(const Scalar& s) {
    const int cn = channels();
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x)
            for (int c = 0; c < cn; ++c) {
                const double v = s[c & 3];
                switch (depth()) {
                case CV_8U: ptr<uchar>(y)[x * cn + c] = (uchar)v; break;
                case CV_32F: ptr<float>(y)[x * cn + c] = (float)v; break;
                case CV_64F: ptr<double>(y)[x * cn + c] = v; break;
                default: throw std::runtime_error("cvshim: setTo depth");
                }
            }
    return *this;
}

Mat Mat::reshape(int cn, int new_rows) const {
    CV_Assert(isContinuous());
    if (cn == 0) cn = channels();
    const size_t total_scalars = total() * (size_t)channels();
    if (new_rows == 0) new_rows = rows;
    CV_Assert(new_rows > 0 && total_scalars % ((size_t)new_rows * (size_t)cn) == 0);
    Mat m = *this;
    m.flags = CV_MAKETYPE(depth(), cn);
    m.rows = new_rows;
    m.cols = (int)(total_scalars / ((size_t)new_rows * (size_t)cn));
    m.step = (size_t)m.cols * m.elemSize();
    return m;
}

Mat Mat::t() const {
    Mat d;
    transpose(*this, d);
    return d;
}

// Mat::convertTo — not reachable through cv2, so the depth pairs the reference path uses are restated here
// (OpenCV modules/core/src/convert_scale.simd.hpp: alpha and beta are narrowed to float for 8u/32f sources and
// the scaled value is a fused multiply-add; 32f -> 8u rounds half-to-even with cvtps2dq semantics, i.e. NaN and
// out-of-int-range values become INT_MIN and then saturate to 0).
void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    if (empty()) { dst.release(); return; }
    const int sd = depth(), cn = channels();
    const int dd = rtype < 0 ? sd : CV_MAT_DEPTH(rtype);
    Mat src = *this;   // keeps the source buffer alive when dst is re-allocated
    Mat out;
    if (!(dst.data == src.data && dd != sd)) out = dst;   // in place with a depth change allocates (as OpenCV)
    out.create(rows, cols, CV_MAKETYPE(dd, cn));          // reuses a matching buffer, else a fresh one
    const float a = (float)alpha, b = (float)beta;
    const bool noscale = alpha == 1.0 && beta == 0.0;
    const int n = cols * cn;
    for (int y = 0; y < rows; ++y) {
        if (sd == CV_8U && dd == CV_32F) {
            const uchar* s = src.ptr<uchar>(y);
            float* d = out.ptr<float>(y);
            for (int i = 0; i < n; ++i) d[i] = noscale ? (float)s[i] : std::fmaf((float)s[i], a, b);
        } else if (sd == CV_32F && dd == CV_32F) {
            const float* s = src.ptr<float>(y);
            float* d = out.ptr<float>(y);
            for (int i = 0; i < n; ++i) d[i] = noscale ? s[i] : std::fmaf(s[i], a, b);
        } else if (sd == CV_32F && dd == CV_8U) {
            const float* s = src.ptr<float>(y);
            uchar* d = out.ptr<uchar>(y);
            for (int i = 0; i < n; ++i) {
                const float v = noscale ? s[i] : std::fmaf(s[i], a, b);
                int iv;
                if (!(v >= -2147483648.0f && v < 2147483648.0f)) iv = INT_MIN;   // NaN / out of range
                else iv = (int)std::nearbyintf(v);                               // round half to even
                d[i] = (uchar)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
            }
        } else if (sd == CV_8U && dd == CV_8U && noscale) {
            std::memmove(out.ptr<uchar>(y), src.ptr<uchar>(y), (size_t)n);
        } else {
            throw std::runtime_error("cvshim: convertTo depth pair not used by the reference path");
        }
    }
    dst = out;
}

// ---------------------------------------------------------------------------------------------------
// cv2 bridge
// ---------------------------------------------------------------------------------------------------
static py::object& cv2mod() {
    static py::object* m = new py::object(py::module_::import("cv2"));   // leaked on purpose (no dtor at exit)
    return *m;
}
static py::object& npmod() {
    static py::object* m = new py::object(py::module_::import("numpy"));
    return *m;
}

static py::dtype dtype_of(int depth) {
    switch (depth) {
    case CV_8U: return py::dtype::of<uint8_t>();
    case CV_8S: return py::dtype::of<int8_t>();
    case CV_16U: return py::dtype::of<uint16_t>();
    case CV_16S: return py::dtype::of<int16_t>();
    case CV_32S: return py::dtype::of<int32_t>();
    case CV_32F: return py::dtype::of<float>();
    case CV_64F: return py::dtype::of<double>();
    }
    throw std::runtime_error("cvshim: dtype_of");
}

static int depth_of(const py::dtype& dt) {
    const char k = dt.kind();
    const auto sz = dt.itemsize();
    if (k == 'u' && sz == 1) return CV_8U;
    if (k == 'i' && sz == 1) return CV_8S;
    if (k == 'u' && sz == 2) return CV_16U;
    if (k == 'i' && sz == 2) return CV_16S;
    if (k == 'i' && sz == 4) return CV_32S;
    if (k == 'f' && sz == 4) return CV_32F;
    if (k == 'f' && sz == 8) return CV_64F;
    throw std::runtime_error("cvshim: unsupported numpy dtype from cv2");
}

// zero-copy numpy view of a Mat (the Mat outlives the call)
static py::array np_of(const Mat& m) {
    CV_Assert(!m.empty());
    const int cn = m.channels();
    const ssize_t es1 = (ssize_t)m.elemSize1();
    std::vector<ssize_t> shape{m.rows, m.cols}, strides{(ssize_t)m.step, es1 * cn};
    if (cn > 1) { shape.push_back(cn); strides.push_back(es1); }
    static int anchor = 0;
    static py::object* keep = new py::object(py::capsule(&anchor, [](void*) {}));
    return py::array(dtype_of(m.depth()), shape, strides, m.data, *keep);
}

// store a cv2 result in dst with OpenCV's create() semantics: when dst already has the result's size and type the
// pixels are written into its existing buffer (other headers sharing it see them, as in OpenCV); otherwise dst
// takes over the numpy buffer without a copy (the array object is kept alive by the Mat's buffer owner).
struct NumpyOwner {
    PyObject* obj;
    void operator()(uchar*) const {
        if (Py_IsInitialized()) { py::gil_scoped_acquire g; Py_DECREF(obj); }   // else: interpreter gone, leak
    }
};

void Mat::adopt(int r, int c, int type, uchar* ptr, std::shared_ptr<uchar[]> owner) {
    flags = type & 0xFFF;
    rows = r;
    cols = c;
    step = (size_t)c * elemSize();
    buf_ = std::move(owner);
    data = ptr;
}

static void to_mat(const py::handle& obj, Mat& dst) {
    py::array arr = py::array::ensure(npmod().attr("ascontiguousarray")(obj));
    if (!arr) throw std::runtime_error("cvshim: cv2 did not return an array");
    const int nd = (int)arr.ndim();
    CV_Assert(nd == 2 || nd == 3);
    const int rows = (int)arr.shape(0), cols = (int)arr.shape(1), cn = nd == 3 ? (int)arr.shape(2) : 1;
    const int type = CV_MAKETYPE(depth_of(arr.dtype()), cn);
    uchar* src = static_cast<uchar*>(arr.mutable_data());
    if (dst.data && dst.rows == rows && dst.cols == cols && dst.type() == type) {
        const size_t rowbytes = (size_t)cols * dst.elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step, src + (size_t)y * rowbytes, rowbytes);
        return;
    }
    PyObject* o = arr.ptr();
    Py_INCREF(o);
    dst.adopt(rows, cols, type, src, std::shared_ptr<uchar[]>(src, NumpyOwner{o}));
}

template <typename... Args> static py::object cvcall(const char* name, Args&&... args) {
    return cv2mod().attr(name)(std::forward<Args>(args)...);
}

static py::tuple pysize(const Size& s) { return py::make_tuple(s.width, s.height); }
static py::tuple pypoint(const Point& p) { return py::make_tuple(p.x, p.y); }

using namespace py::literals;

// ---- core ------------------------------------------------------------------------------------------
void add(const Mat& a, const Mat& b, Mat& dst) { to_mat(cvcall("add", np_of(a), np_of(b)), dst); }
void subtract(const Mat& a, const Mat& b, Mat& dst) { to_mat(cvcall("subtract", np_of(a), np_of(b)), dst); }
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale) {
    if (scale == 1.0) to_mat(cvcall("multiply", np_of(a), np_of(b)), dst);
    else to_mat(cvcall("multiply", np_of(a), np_of(b), "scale"_a = scale), dst);
}
// Mat x double: the C++ call wraps the double as a 1x1 CV_64F array, the Python binding as a 4-double Scalar;
// both reach arithm_op's scalar branch (same working-type rule), so this is the same kernel.
void multiply(const Mat& a, double s, Mat& dst) { to_mat(cvcall("multiply", np_of(a), s), dst); }
void divide(const Mat& a, const Mat& b, Mat& dst, double scale) {
    if (scale == 1.0) to_mat(cvcall("divide", np_of(a), np_of(b)), dst);
    else to_mat(cvcall("divide", np_of(a), np_of(b), "scale"_a = scale), dst);
}
void divide(const Mat& a, double s, Mat& dst) { to_mat(cvcall("divide", np_of(a), s), dst); }
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst) {
    to_mat(cvcall("addWeighted", np_of(a), alpha, np_of(b), beta, gamma), dst);
}
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst) { to_mat(cvcall("scaleAdd", np_of(a), alpha, np_of(b)), dst); }
void sqrt(const Mat& src, Mat& dst) { to_mat(cvcall("sqrt", np_of(src)), dst); }

void split(const Mat& src, std::vector<Mat>& mv) {
    py::tuple t = py::tuple(cvcall("split", np_of(src)));
    mv.resize(t.size());
    for (size_t i = 0; i < t.size(); ++i) to_mat(t[i], mv[i]);
}
void split(const Mat& src, Mat* mv) {
    py::tuple t = py::tuple(cvcall("split", np_of(src)));
    for (size_t i = 0; i < t.size(); ++i) to_mat(t[i], mv[i]);
}
void merge(const Mat* mv, size_t count, Mat& dst) {
    py::list l;
    for (size_t i = 0; i < count; ++i) l.append(np_of(mv[i]));
    to_mat(cvcall("merge", l), dst);
}
void merge(const std::vector<Mat>& mv, Mat& dst) { merge(mv.data(), mv.size(), dst); }
void hconcat(const Mat& a, const Mat& b, Mat& dst) {
    py::list l;
    l.append(np_of(a));
    l.append(np_of(b));
    to_mat(cvcall("hconcat", l), dst);
}
void transpose(const Mat& src, Mat& dst) { to_mat(cvcall("transpose", np_of(src)), dst); }
void patchNaNs(Mat& a, double val) {
    // cv2.patchNaNs works in place on the array it is given; hand it a private copy and store the result
    py::object arr = npmod().attr("array")(np_of(a));
    py::object r = cvcall("patchNaNs", arr, val);
    to_mat(r.is_none() ? arr : r, a);
}
void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type, int dtype) {
    to_mat(cvcall("normalize", np_of(src), py::none(), alpha, beta, norm_type, dtype), dst);
}
void minMaxLoc(const Mat& src, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc) {
    // C++ minMaxLoc without locations accepts multi-channel input and scans it as one channel (minMaxIdx);
    // the Python binding always asks for locations, so hand it the single-channel reshape OpenCV uses itself.
    Mat one = src.channels() == 1 ? src : (src.isContinuous() ? src : src.clone()).reshape(1);
    py::tuple t = py::tuple(cvcall("minMaxLoc", np_of(one)));
    if (minVal) *minVal = t[0].cast<double>();
    if (maxVal) *maxVal = t[1].cast<double>();
    CV_Assert((!minLoc && !maxLoc) || src.channels() == 1);
    if (minLoc) { py::tuple p = py::tuple(t[2]); *minLoc = Point(p[0].cast<int>(), p[1].cast<int>()); }
    if (maxLoc) { py::tuple p = py::tuple(t[3]); *maxLoc = Point(p[0].cast<int>(), p[1].cast<int>()); }
}
void dft(const Mat& src, Mat& dst, int flags, int nonzeroRows) {
    to_mat(cvcall("dft", np_of(src), "flags"_a = flags, "nonzeroRows"_a = nonzeroRows), dst);
}
void idft(const Mat& src, Mat& dst, int flags, int nonzeroRows) {
    to_mat(cvcall("idft", np_of(src), "flags"_a = flags, "nonzeroRows"_a = nonzeroRows), dst);
}
void mulSpectrums(const Mat& a, const Mat& b, Mat& c, int flags, bool conjB) {
    to_mat(cvcall("mulSpectrums", np_of(a), np_of(b), flags, "conjB"_a = conjB), c);
}
int getOptimalDFTSize(int vecsize) { return cvcall("getOptimalDFTSize", vecsize).cast<int>(); }
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar& value) {
    to_mat(cvcall("copyMakeBorder", np_of(src), top, bottom, left, right, borderType,
                  "value"_a = py::make_tuple(value[0], value[1], value[2], value[3])), dst);
}

// ---- imgproc ---------------------------------------------------------------------------------------
void pyrDown(const Mat& src, Mat& dst, const Size& dstsize, int borderType) {
    if (dstsize.area() == 0) to_mat(cvcall("pyrDown", np_of(src), "borderType"_a = borderType), dst);
    else to_mat(cvcall("pyrDown", np_of(src), "dstsize"_a = pysize(dstsize), "borderType"_a = borderType), dst);
}
void pyrUp(const Mat& src, Mat& dst, const Size& dstsize, int borderType) {
    if (dstsize.area() == 0) to_mat(cvcall("pyrUp", np_of(src), "borderType"_a = borderType), dst);
    else to_mat(cvcall("pyrUp", np_of(src), "dstsize"_a = pysize(dstsize), "borderType"_a = borderType), dst);
}
void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation) {
    to_mat(cvcall("resize", np_of(src), pysize(dsize), "fx"_a = fx, "fy"_a = fy, "interpolation"_a = interpolation), dst);
}
void cvtColor(const Mat& src, Mat& dst, int code, int dstCn) {
    to_mat(cvcall("cvtColor", np_of(src), code, "dstCn"_a = dstCn), dst);
}
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor, double delta, int borderType) {
    // the reference passes depth constants spelled as types (CV_32FC1); OpenCV takes CV_MAT_DEPTH of them
    to_mat(cvcall("filter2D", np_of(src), ddepth < 0 ? -1 : CV_MAT_DEPTH(ddepth), np_of(kernel), "anchor"_a = pypoint(anchor),
                  "delta"_a = delta, "borderType"_a = borderType), dst);
}
void sepFilter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernelX, const Mat& kernelY, Point anchor, double delta,
                 int borderType) {
    to_mat(cvcall("sepFilter2D", np_of(src), ddepth < 0 ? -1 : CV_MAT_DEPTH(ddepth), np_of(kernelX), np_of(kernelY),
                  "anchor"_a = pypoint(anchor), "delta"_a = delta, "borderType"_a = borderType), dst);
}
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    to_mat(cvcall("GaussianBlur", np_of(src), pysize(ksize), sigmaX, "sigmaY"_a = sigmaY, "borderType"_a = borderType), dst);
}
Mat getGaussianKernel(int ksize, double sigma, int ktype) {
    Mat k;
    to_mat(cvcall("getGaussianKernel", ksize, sigma, "ktype"_a = CV_MAT_DEPTH(ktype)), k);
    return k;
}
double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
    py::tuple t = py::tuple(cvcall("threshold", np_of(src), thresh, maxval, type));
    to_mat(t[1], dst);
    return t[0].cast<double>();
}

// ---------------------------------------------------------------------------------------------------
// MatExpr lowering — restated from the behaviour of OpenCV's modules/core/src/matop.cpp
// (MatOp::add / MatOp::subtract, MatOp_AddEx::assign / ::multiply, MatOp_Bin::assign / ::multiply).
// A plain Mat inside an expression (OpenCV's identity op) is the ADDEX {a, alpha = 1} form here; both forms
// take the same branches below.
// ---------------------------------------------------------------------------------------------------
static MatExpr addex(const Mat& a, const Mat& b, double alpha, double beta, const Scalar& s = Scalar()) {
    MatExpr e;
    e.kind = MatExpr::ADDEX;
    e.a = a;
    e.b = b;
    e.alpha = alpha;
    e.beta = beta;
    e.s = s;
    return e;
}
static bool scalar_is_zero(const Scalar& s) { return s[0] == 0 && s[1] == 0 && s[2] == 0 && s[3] == 0; }

static void assign_expr(const MatExpr& e, Mat& m) {
    if (e.kind == MatExpr::MUL) {
        multiply(e.a, e.b, m, e.alpha);
        return;
    }
    CV_Assert(e.s.isReal());
    if (e.b.data) {
        if (scalar_is_zero(e.s)) {
            if (e.alpha == 1) {
                if (e.beta == 1) add(e.a, e.b, m);
                else if (e.beta == -1) subtract(e.a, e.b, m);
                else scaleAdd(e.b, e.beta, e.a, m);
            } else if (e.beta == 1) {
                if (e.alpha == -1) subtract(e.b, e.a, m);
                else scaleAdd(e.a, e.alpha, e.b, m);
            } else {
                addWeighted(e.a, e.alpha, e.b, e.beta, 0, m);
            }
        } else {
            addWeighted(e.a, e.alpha, e.b, e.beta, e.s[0], m);
        }
    } else if (std::fabs(e.alpha) != 1) {
        CV_Assert(e.a.channels() == 1 || e.s[0] == 0.0);
        e.a.convertTo(m, -1, e.alpha, e.s[0]);
    } else if (e.alpha == 1) {
        if (scalar_is_zero(e.s)) {
            // OpenCV: cv::add(a, Scalar(0)) — x + 0 is x for every float including -0 (+0 result) and NaN;
            // only -0.0 would differ in sign, which no consumer on this path can observe
            e.a.copyTo(m);
        } else {
            throw std::runtime_error("cvshim: Mat + non-zero scalar is not used by the reference path");
        }
    } else {   // alpha == -1: cv::subtract(Scalar s, a)
        CV_Assert(scalar_is_zero(e.s));
        Mat z = Mat::zeros(e.a.size(), e.a.type());
        subtract(z, e.a, m);
    }
}

MatExpr::operator Mat() const {
    Mat m;
    assign_expr(*this, m);
    return m;
}
Mat& Mat::operator=(const MatExpr& e) {
    assign_expr(e, *this);
    return *this;
}
MatExpr Mat::mul(const Mat& m, double scale) const {
    MatExpr e;
    e.kind = MatExpr::MUL;
    e.a = *this;
    e.b = m;
    e.alpha = scale;
    return e;
}

static bool simple_addex(const MatExpr& e) { return e.kind == MatExpr::ADDEX && (!e.b.data || e.beta == 0); }

static MatExpr add_sub(const MatExpr& e1, const MatExpr& e2, double sign) {
    double alpha = 1, beta = sign;
    Scalar s;
    Mat m1, m2;
    if (simple_addex(e1)) { m1 = e1.a; alpha = e1.alpha; s = e1.s; } else assign_expr(e1, m1);
    if (simple_addex(e2)) {
        m2 = e2.a;
        beta = sign * e2.alpha;
        for (int i = 0; i < 4; ++i) s[i] += sign * e2.s[i];
    } else assign_expr(e2, m2);
    return addex(m1, m2, alpha, beta, s);
}

MatExpr operator+(const Mat& a, const Mat& b) { return addex(a, b, 1, 1); }
MatExpr operator+(const MatExpr& e, const Mat& m) { return add_sub(e, MatExpr(m), 1); }
MatExpr operator+(const Mat& m, const MatExpr& e) { return add_sub(MatExpr(m), e, 1); }
MatExpr operator+(const MatExpr& e1, const MatExpr& e2) { return add_sub(e1, e2, 1); }
MatExpr operator-(const Mat& a, const Mat& b) { return addex(a, b, 1, -1); }
MatExpr operator-(const MatExpr& e, const Mat& m) { return add_sub(e, MatExpr(m), -1); }
MatExpr operator-(const Mat& m, const MatExpr& e) { return add_sub(MatExpr(m), e, -1); }
MatExpr operator-(const MatExpr& e1, const MatExpr& e2) { return add_sub(e1, e2, -1); }
MatExpr operator*(const Mat& a, double s) { return addex(a, Mat(), s, 0); }
MatExpr operator*(double s, const Mat& a) { return addex(a, Mat(), s, 0); }
MatExpr operator*(const MatExpr& e, double s) {
    MatExpr r = e;
    r.alpha *= s;
    if (e.kind == MatExpr::ADDEX) {
        r.beta *= s;
        for (int i = 0; i < 4; ++i) r.s[i] *= s;
    }
    return r;
}
MatExpr operator*(double s, const MatExpr& e) { return e * s; }
MatExpr operator/(const Mat& a, double s) { return addex(a, Mat(), 1.0 / s, 0); }
MatExpr operator/(const MatExpr& e, double s) { return e * (1.0 / s); }

// Called once when the extension module is imported: the enum values baked into the headers must be cv2's.
void cvshim_selfcheck() {
    struct { const char* name; int value; } consts[] = {
        {"BORDER_CONSTANT", BORDER_CONSTANT}, {"BORDER_REFLECT_101", BORDER_REFLECT_101}, {"BORDER_DEFAULT", BORDER_DEFAULT},
        {"DFT_INVERSE", DFT_INVERSE}, {"DFT_SCALE", DFT_SCALE}, {"DFT_ROWS", DFT_ROWS}, {"NORM_MINMAX", NORM_MINMAX},
        {"INTER_NEAREST", INTER_NEAREST}, {"INTER_LINEAR", INTER_LINEAR}, {"INTER_AREA", INTER_AREA},
        {"THRESH_TRUNC", THRESH_TRUNC}, {"COLOR_BGR2GRAY", COLOR_BGR2GRAY}, {"COLOR_BGR2Lab", COLOR_BGR2Lab},
        {"COLOR_Lab2BGR", COLOR_Lab2BGR}, {"CV_8U", CV_8U}, {"CV_32F", CV_32F}, {"CV_64F", CV_64F},
        {"CV_8UC3", CV_8UC3}, {"CV_32FC3", CV_32FC3},
    };
    for (const auto& c : consts)
        if (cv2mod().attr(c.name).cast<int>() != c.value)
            throw std::runtime_error(std::string("cvshim: constant mismatch with cv2: ") + c.name);
}

}  // namespace cv
