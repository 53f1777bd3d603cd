// TEST STUB: the few cv::Mat members the adapter touches, so it can be compile-checked without OpenCV.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    int type_ = CV_8UC3;
    size_t step = 0;
    uint8_t* data = nullptr;
    std::vector<uint8_t> store;
    Mat() = default;
    Mat(int r, int c, int t) : rows(r), cols(c), type_(t) {
        step = (size_t)c * channels();
        store.resize(step * (size_t)r);
        data = store.data();
    }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), type_(o.type_), step(o.step), store(o.store) { data = store.empty() ? o.data : store.data(); }
    Mat(Mat&& o) noexcept : rows(o.rows), cols(o.cols), type_(o.type_), step(o.step), store(std::move(o.store)) { data = store.empty() ? o.data : store.data(); }
    Mat& operator=(Mat&& o) noexcept { rows = o.rows; cols = o.cols; type_ = o.type_; step = o.step; store = std::move(o.store); data = store.empty() ? o.data : store.data(); return *this; }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; type_ = o.type_; step = o.step; store = o.store; data = store.empty() ? o.data : store.data(); return *this; }
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    int type() const { return type_; }
};
}  // namespace cv
