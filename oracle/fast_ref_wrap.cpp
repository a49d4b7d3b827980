/*
 * fast_ref_wrap.cpp -- TEST INFRASTRUCTURE. Thin extern "C" shim over the REFERENCE's own FAST library
 * (Thirdparty/fast/src/*.cpp, compiled from where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libfast_ref.so; no reference source is copied into this repo). Used to pin the oracle's
 * closed-form FAST-10 and as the "reference" CPU timing for the FAST stage.
 *   ref: Thirdparty/fast/include/fast/fast.h:20-29
 */
#include <fast/fast.h>
#include <cstddef>
#include <cstdint>
#include <vector>

extern "C" {

// sse2 != 0 -> fast_corner_detect_10_sse2 (what Feature_detector::detect calls on x86, ref: src/Feature_detection.cpp:80-82)
int fastref_detect10(const uint8_t* img, int w, int h, int stride, int barrier, int sse2, int16_t* xy, int cap)
{
    std::vector<fast::fast_xy> c;
    if (sse2) fast::fast_corner_detect_10_sse2(img, w, h, stride, (short)barrier, c);
    else      fast::fast_corner_detect_10(img, w, h, stride, (short)barrier, c);
    const int n = (int)c.size();
    for (int i = 0; i < n && i < cap; ++i) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; }
    return n;
}

void fastref_score10(const uint8_t* img, int stride, const int16_t* xy, int n, int threshold, int* scores)
{
    std::vector<fast::fast_xy> c;
    c.reserve(n);
    for (int i = 0; i < n; ++i) c.emplace_back(xy[2 * i], xy[2 * i + 1]);
    std::vector<int> s;
    fast::fast_corner_score_10(img, stride, c, threshold, s);
    for (int i = 0; i < n; ++i) scores[i] = s[i];
}

int fastref_nonmax(const int16_t* xy, const int* scores, int n, int* keep)
{
    std::vector<fast::fast_xy> c;
    c.reserve(n);
    for (int i = 0; i < n; ++i) c.emplace_back(xy[2 * i], xy[2 * i + 1]);
    std::vector<int> s(scores, scores + n), k;
    fast::fast_nonmax_3x3(c, s, k);
    for (std::size_t i = 0; i < k.size(); ++i) keep[i] = k[i];
    return (int)k.size();
}

}  // extern "C"
