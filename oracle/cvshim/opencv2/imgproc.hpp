// cvshim — TEST INFRASTRUCTURE ONLY.  imgproc slice of the OpenCV-shaped facade (see core.hpp); every function
// below is forwarded to the real OpenCV kernel of the same name through cv2 (cvshim.cpp).
#pragma once
#include "opencv2/core.hpp"

namespace cv {

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum ThresholdTypes { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3 };
enum ColorConversionCodes { COLOR_BGR2GRAY = 6, COLOR_BGR2Lab = 44, COLOR_Lab2BGR = 56 };

void pyrDown(const Mat& src, Mat& dst, const Size& dstsize = Size(), int borderType = BORDER_DEFAULT);
void pyrUp(const Mat& src, Mat& dst, const Size& dstsize = Size(), int borderType = BORDER_DEFAULT);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void cvtColor(const Mat& src, Mat& dst, int code, int dstCn = 0);
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0,
              int borderType = BORDER_DEFAULT);
void sepFilter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernelX, const Mat& kernelY, Point anchor = Point(-1, -1),
                 double delta = 0, int borderType = BORDER_DEFAULT);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
Mat getGaussianKernel(int ksize, double sigma, int ktype = CV_64F);
double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type);

}  // namespace cv
