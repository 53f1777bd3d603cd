// Compile + link + run check of the reference-side adapter against stub reference headers.
// Without a GPU the constructor must throw (no CPU fallback); with one it magnifies a frame.
#include <cstdio>
#include <cstring>
#include "MagnificationProcessorB200.hpp"
int main() {
    using namespace livim;
    try {
        MagnificationProcessorB200 proc(0);
        ProcessorConfig cfg;
        cfg.magnification.mode = MagnificationMode::Laplace;
        cfg.magnification.amplification = 20; cfg.magnification.coWavelength = 500;
        cfg.magnification.coLow = 0.08; cfg.magnification.coHigh = 0.46; cfg.magnification.levels = 4;
        auto f = std::make_shared<Frame>();
        f->image = cv::Mat(48, 64, CV_8UC3);
        std::memset(f->image.data, 90, f->image.step * (size_t)f->image.rows);
        f->seq = 42;
        FrameRef out = proc.process(f, cfg);
        if (out == f || out->seq != 42 || out->image.data == f->image.data) { std::puts("FAIL"); return 1; }
        cfg.magnification.mode = MagnificationMode::None;
        if (proc.process(f, cfg) != f) { std::puts("FAIL identity"); return 1; }
        proc.reset();
        std::puts("OK gpu");
        return 0;
    } catch (const std::runtime_error& e) {
        std::printf("THROWN %s\n", e.what());
        return 3;
    }
}
