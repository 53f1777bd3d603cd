// TEST STUB with the fields of reference src/core/Frame.hpp:16-30.
#pragma once
#include <cstdint>
#include <memory>
#include <opencv2/core.hpp>
namespace livim {
enum class PixelFormat { BGR8, Gray8 };
struct Frame {
    std::uint64_t seq = 0;
    std::int64_t ptsUs = 0;
    double captureTs = 0;
    int width = 0, height = 0;
    PixelFormat format = PixelFormat::BGR8;
    cv::Mat image;
};
using FrameRef = std::shared_ptr<const Frame>;
using MutableFrameRef = std::shared_ptr<Frame>;
}  // namespace livim
