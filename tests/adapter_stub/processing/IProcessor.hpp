// TEST STUB with the declarations of reference src/processing/IProcessor.hpp:10-60 (same names/fields).
#pragma once
#include "core/Frame.hpp"
namespace livim {
enum class MagnificationMode { Laplace, Phase, Color, None };
struct MagnificationParams {
    MagnificationMode mode = MagnificationMode::Laplace;
    double amplification = 0.0, coWavelength = 0.0, coLow = 0.0, coHigh = 0.0, chromAttenuation = 0.0;
    int levels = 4;
    double framerate = 30.0;
};
struct PreprocessParams {
    int downscale = 1;
    bool roiEnabled = false;
    float roiX = 0.0f, roiY = 0.0f, roiW = 1.0f, roiH = 1.0f;
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
