// ORACLE - TEST INFRASTRUCTURE ONLY.  Stand-in for <opencv2/opencv.hpp> (absent from the image): thirdparty/LidarIris/LidarIris.h only
// DECLARES members with these types; the loop detector that uses them is not built (see oracle/src/ref_estimator_wrap.cpp).
#pragma once
namespace cv {
struct Mat {};
struct Mat1b : Mat {};
struct Mat1f : Mat {};
struct Mat2f : Mat {};
}  // namespace cv
