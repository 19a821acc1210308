// mat.hpp - the few pieces of cv::Mat the USAC plugin surface touches, as a row-major float32 shim, so that the host layer
// builds without OpenCV (absent from this image). Define USAC_WITH_OPENCV to use the real headers instead.
#pragma once
#ifdef USAC_WITH_OPENCV
#include <opencv2/core.hpp>
#else
#include <cassert>
#include <cstring>
#include <memory>
#include <ostream>
#include <vector>

namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;              // float32, row-major, continuous (what the reference assumes, ransac.hpp:41)
    Mat() = default;
    Mat(int r, int c) : rows(r), cols(c), store_(std::make_shared<std::vector<float>>((size_t)r * c, 0.f)) { data = bytes(); }
    Mat(int r, int c, const float* borrowed) : rows(r), cols(c), data((unsigned char*)borrowed) {}   // header over caller memory
    bool empty() const { return data == nullptr || rows * cols == 0; }
    Mat clone() const {
        Mat m(rows, cols);
        if (!empty()) std::memcpy(m.data, data, sizeof(float) * rows * cols);
        return m;
    }
    const Mat& getMat() const { return *this; }                        // cv::InputArray::getMat()
    float& at(int r, int c) { return ((float*)data)[(size_t)r * cols + c]; }
    float at(int r, int c) const { return ((const float*)data)[(size_t)r * cols + c]; }
    const float* ptr() const { return (const float*)data; }
    float* ptr() { return (float*)data; }
    static Mat eye(int n) { Mat m(n, n); for (int i = 0; i < n; i++) m.at(i, i) = 1.f; return m; }
private:
    std::shared_ptr<std::vector<float>> store_;
    unsigned char* bytes() { return (unsigned char*)store_->data(); }
};
typedef const Mat& InputArray;
inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    os << "[";
    for (int r = 0; r < m.rows; r++) {
        for (int c = 0; c < m.cols; c++) os << m.at(r, c) << (c + 1 < m.cols ? ", " : "");
        os << (r + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}
}   // namespace cv
#endif
