// TEST STUB — NOT reference code.  A minimal stand-in for the three reference/third-party headers the adapter
// includes ("opencv2/core.hpp", "core/Frame.hpp", "processing/IProcessor.hpp"), so that
// adapter/MagnificationProcessorB200.hpp can be compile-checked and run without OpenCV or the reference tree.
// Only the members the adapter touches are modelled; field names follow the reference interface because the
// adapter is written against it (reference src/core/Frame.hpp:16-30, src/processing/IProcessor.hpp:10-60).
// __graft_entry__.build() generates forwarding headers with the expected include names under _gen/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16

namespace cv {
// owning 8-bit image with the cv::Mat members used by the adapter: rows, cols, step, data, type(), channels(), empty()
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    std::vector<uint8_t> store;

    Mat() = default;
    Mat(int r, int c, int t) : rows(r), cols(c), type_(t) { allocate(); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), step(o.step), store(o.store), type_(o.type_) { rebind(o); }
    Mat(Mat&& o) noexcept : rows(o.rows), cols(o.cols), step(o.step), store(std::move(o.store)), type_(o.type_) { rebind(o); }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; step = o.step; type_ = o.type_; store = o.store; rebind(o); return *this; }
    Mat& operator=(Mat&& o) noexcept { rows = o.rows; cols = o.cols; step = o.step; type_ = o.type_; store = std::move(o.store); rebind(o); return *this; }

    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    int type() const { return type_; }

private:
    int type_ = CV_8UC3;
    void allocate() { step = (size_t)cols * channels(); store.assign(step * (size_t)rows, 0); data = store.data(); }
    void rebind(const Mat& o) { data = store.empty() ? o.data : store.data(); }
};
}  // namespace cv

namespace livim {

enum class PixelFormat { BGR8, Gray8 };

struct Frame {   // metadata + image, as consumed by the adapter
    std::uint64_t seq = 0;
    std::int64_t ptsUs = 0;
    double captureTs = 0;
    int width = 0, height = 0;
    PixelFormat format = PixelFormat::BGR8;
    cv::Mat image;
};
using FrameRef = std::shared_ptr<const Frame>;

enum class MagnificationMode { Laplace, Phase, Color, None };

struct MagnificationParams {
    MagnificationMode mode = MagnificationMode::Laplace;
    double amplification = 0, coWavelength = 0, coLow = 0, coHigh = 0, chromAttenuation = 0;
    int levels = 4;
    double framerate = 30;
};

struct PreprocessParams {
    int downscale = 1;
    bool roiEnabled = false;
    float roiX = 0, roiY = 0, roiW = 1, roiH = 1;
};

struct ProcessorConfig {
    bool grayscale = false;
    PreprocessParams preprocess;
    MagnificationParams magnification;
};

class IProcessor {
public:
    virtual ~IProcessor() = default;
    virtual FrameRef process(const FrameRef& in, const ProcessorConfig& cfg) = 0;
    virtual void reset() {}
};

}  // namespace livim
