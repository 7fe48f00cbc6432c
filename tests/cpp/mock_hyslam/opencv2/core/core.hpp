// TEST DOUBLE of the few OpenCV types hySLAM's feature interfaces mention (cv::Mat, cv::KeyPoint, cv::InputArray).
// This image has no OpenCV C++ headers; the double lets include/hyorb_hyslam.hpp be compiled and exercised here.  Field
// names, layouts and member signatures follow OpenCV 3.x so that the shim source is the one a hySLAM build would use.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint {            // 28 bytes, same member order as OpenCV's
    Point2f pt; float size = 0; float angle = -1; float response = 0; int octave = 0; int class_id = -1;
};
class Mat {
public:
    int rows = 0, cols = 0; unsigned char *data = nullptr; size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext, size_t stp = 0) : rows(r), cols(c), data((unsigned char *)ext), type_(type) { step = stp ? stp : (size_t)c * esz(); }
    void create(int r, int c, int type) { rows = r; cols = c; type_ = type; step = (size_t)c * esz(); store_ = std::make_shared<std::vector<unsigned char>>(step * (size_t)r); data = store_->data(); }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * esz(); }
    Mat clone() const { Mat m; if (empty()) return m; m.create(rows, cols, type_); for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz()); return m; }
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
private:
    size_t esz() const { return ((type_ & 7) == CV_32F ? 4 : 1) * (size_t)channels(); }
    int type_ = CV_8U; std::shared_ptr<std::vector<unsigned char>> store_;
};
class _InputArray {           // cv::InputArray = const _InputArray&
public:
    _InputArray() {}
    _InputArray(const Mat &m) : m_(m) {}
    Mat getMat() const { return m_; }
    bool empty() const { return m_.empty(); }
private:
    Mat m_;
};
typedef const _InputArray &InputArray;
}  // namespace cv
