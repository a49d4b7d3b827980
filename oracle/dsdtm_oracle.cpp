/*
 * dsdtm_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see dsdtm_oracle.h header).
 *
 * Dependency-free CPU restatement of DSDTM's tracking front end. Build with
 *   g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared   (oracle/Makefile)
 * -ffp-contract=off mirrors the reference's FMA-free x86-64 build (ref: CMakeLists.txt:4-8).
 * Every function cites the reference lines it follows ("ref:" = below /root/reference).
 */
#include "dsdtm_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// Sophus (non-templated 1.0) / Eigen quaternion semantics, SURVEY App. B.3. Not in /root/reference.
// ------------------------------------------------------------------------------------------
struct Quat { double w, x, y, z; };
struct Se3 { Quat q; double t[3]; };

inline Quat qmul(const Quat& a, const Quat& b)  // Eigen::Quaternion product
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
inline void qnormalize(Quat& q)
{
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
inline void qrot(const Quat& q, const double v[3], double out[3])  // Eigen QuaternionBase::_transformVector
{
    double uv[3] = { q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    double c[3] = { q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0] };
    out[0] = v[0] + q.w * uv[0] + c[0];
    out[1] = v[1] + q.w * uv[1] + c[1];
    out[2] = v[2] + q.w * uv[2] + c[2];
}
inline Se3 se3_from(const double p[7]) { Se3 s; s.q = { p[0], p[1], p[2], p[3] }; s.t[0] = p[4]; s.t[1] = p[5]; s.t[2] = p[6]; return s; }
inline void se3_to(const Se3& s, double p[7]) { p[0] = s.q.w; p[1] = s.q.x; p[2] = s.q.y; p[3] = s.q.z; p[4] = s.t[0]; p[5] = s.t[1]; p[6] = s.t[2]; }
inline void se3_act(const Se3& a, const double p[3], double out[3])
{
    double r[3]; qrot(a.q, p, r);
    out[0] = r[0] + a.t[0]; out[1] = r[1] + a.t[1]; out[2] = r[2] + a.t[2];
}
inline Se3 se3_mul(const Se3& a, const Se3& b)  // SE3::operator*= : t += R*t_b ; q *= q_b ; normalize
{
    Se3 r;
    double rt[3]; qrot(a.q, b.t, rt);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    r.q = qmul(a.q, b.q);
    qnormalize(r.q);
    return r;
}
inline Se3 se3_inv(const Se3& a)  // ret.so3 = conj ; ret.t = ret.so3 * (t * -1)
{
    Se3 r;
    r.q = { a.q.w, -a.q.x, -a.q.y, -a.q.z };
    double nt[3] = { a.t[0] * -1., a.t[1] * -1., a.t[2] * -1. };
    qrot(r.q, nt, r.t);
    return r;
}
inline void quat_to_R(const Quat& q, double R[9])  // Eigen toRotationMatrix
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
Se3 se3_exp(const double x[6])  // SE3::exp / SO3::expAndTheta, translation first (upsilon), rotation last (omega)
{
    const double SMALL_EPS = 1e-10;
    const double* ups = x;
    const double* om = x + 3;
    const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double half = 0.5 * theta;
    double imag;
    const double real = std::cos(half);
    if (theta < SMALL_EPS) {
        const double t2 = theta * theta, t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = std::sin(half) / theta;
    }
    Se3 r;
    r.q = { real, imag * om[0], imag * om[1], imag * om[2] };
    // Eigen Quaternion ctor does not normalise; Sophus SO3(Quaternion) ctor does.
    qnormalize(r.q);
    // Omega = hat(omega)
    const double O[9] = { 0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0 };
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += O[3 * i + k] * O[3 * k + j];
            O2[3 * i + j] = s;
        }
    double V[9];
    if (theta < SMALL_EPS) {
        quat_to_R(r.q, V);
    } else {
        const double t2 = theta * theta;
        const double a = (1 - std::cos(theta)) / t2;
        const double b = (theta - std::sin(theta)) / (t2 * theta);
        for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    }
    for (int i = 0; i < 3; ++i) r.t[i] = V[3 * i] * ups[0] + V[3 * i + 1] * ups[1] + V[3 * i + 2] * ups[2];
    return r;
}

// ------------------------------------------------------------------------------------------
// Eigen LDLT<Matrix6d>::compute + solve (pivoted, lower; App. B.4). Not in /root/reference.
// Returns x = H^-1 b with Eigen's "zero pivot -> zero component" pseudo-inverse rule.
// ------------------------------------------------------------------------------------------
void ldlt6_solve(const double Hin[36], const double bin[6], double x[6])
{
    const int n = 6;
    double A[36];
    std::memcpy(A, Hin, sizeof(A));  // row-major, only lower triangle is referenced
    int tr[6];
    auto L = [&](int i, int j) -> double& { return A[i * n + j]; };
    bool zero_matrix = false;
    for (int k = 0; k < n; ++k) {
        int big = k; double bigv = std::fabs(L(k, k));
        for (int i = k + 1; i < n; ++i) { double v = std::fabs(L(i, i)); if (v > bigv) { bigv = v; big = i; } }
        tr[k] = big;
        if (k != big) {
            // symmetric swap of rows/cols k and big in the lower triangle
            for (int j = 0; j < k; ++j) std::swap(L(k, j), L(big, j));
            for (int i = big + 1; i < n; ++i) std::swap(L(i, k), L(i, big));
            std::swap(L(k, k), L(big, big));
            for (int i = k + 1; i < big; ++i) std::swap(L(i, k), L(big, i));
        }
        const int rs = n - k - 1;
        if (k > 0) {
            double temp[6];
            for (int j = 0; j < k; ++j) temp[j] = L(j, j) * L(k, j);
            double s = 0; for (int j = 0; j < k; ++j) s += L(k, j) * temp[j];
            L(k, k) -= s;
            for (int i = k + 1; i < n; ++i) {
                double s2 = 0; for (int j = 0; j < k; ++j) s2 += L(i, j) * temp[j];
                L(i, k) -= s2;
            }
        }
        const double akk = L(k, k);
        const bool valid = std::fabs(akk) > 0.0;
        if (k == 0 && !valid) {
            for (int j = 0; j < n; ++j) tr[j] = j;
            zero_matrix = true;
            break;
        }
        if (rs > 0 && valid) for (int i = k + 1; i < n; ++i) L(i, k) /= akk;
    }
    (void)zero_matrix;
    double y[6];
    for (int i = 0; i < n; ++i) y[i] = bin[i];
    for (int k = 0; k < n; ++k) if (tr[k] != k) std::swap(y[k], y[tr[k]]);   // P b
    for (int i = 0; i < n; ++i) { for (int j = 0; j < i; ++j) y[i] -= L(i, j) * y[j]; }  // L^-1
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int i = 0; i < n; ++i) { if (std::fabs(L(i, i)) > tol) y[i] /= L(i, i); else y[i] = 0; }
    for (int i = n - 1; i >= 0; --i) { for (int j = i + 1; j < n; ++j) y[i] -= L(j, i) * y[j]; }  // L^-T
    for (int k = n - 1; k >= 0; --k) if (tr[k] != k) std::swap(y[k], y[tr[k]]);  // P^T
    for (int i = 0; i < n; ++i) x[i] = y[i];
}

inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

// FAST ring, ref: Thirdparty/fast/src/fast_10.cpp:17-34
const int RING_DX[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
const int RING_DY[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

inline bool has_arc10(unsigned m)  // 10 contiguous set bits in a cyclic 16-bit mask
{
    unsigned d = m | (m << 16);
    unsigned r = d;
    r &= r >> 1;   // runs of 2
    r &= r >> 2;   // runs of 4
    r &= r >> 4;   // runs of 8
    r &= d >> 8;   // runs of 9
    r &= d >> 9;   // runs of 10
    return (r & 0xFFFFu) != 0;
}

}  // namespace

extern "C" {

int orc_cvround(double v) { return (int)std::nearbyint(v); }  // cvRound: round-half-to-even (lrint)

// ------------------------------------------------------------------------------------------
// B.1 pyrDown. ref call site: src/Frame.cpp:79. Separable [1 4 6 4 1], reflect-101, (sum+128)>>8.
// ------------------------------------------------------------------------------------------
void orc_pyrdown_u8(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst)
{
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    static const int K[5] = { 1, 4, 6, 4, 1 };
    std::vector<int> rowbuf((size_t)h * dw);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < dw; ++x) {
            int s = 0;
            for (int j = -2; j <= 2; ++j) s += K[j + 2] * src[(size_t)y * src_stride + reflect101(2 * x + j, w)];
            rowbuf[(size_t)y * dw + x] = s;
        }
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int s = 0;
            for (int i = -2; i <= 2; ++i) s += K[i + 2] * rowbuf[(size_t)reflect101(2 * y + i, h) * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((s + 128) >> 8);
        }
}

// ref: src/Frame.cpp:74-81
void orc_pyramid(const uint8_t* img, int w, int h, int levels, uint8_t* out, int* offs, int* ws, int* hs)
{
    int off = 0;
    for (int l = 0; l < levels; ++l) {
        ws[l] = (l == 0) ? w : (ws[l - 1] + 1) / 2;
        hs[l] = (l == 0) ? h : (hs[l - 1] + 1) / 2;
        offs[l] = off;
        off += ws[l] * hs[l];
    }
    std::memcpy(out, img, (size_t)w * h);
    for (int l = 1; l < levels; ++l) orc_pyrdown_u8(out + offs[l - 1], ws[l - 1], hs[l - 1], ws[l - 1], out + offs[l]);
}

// ------------------------------------------------------------------------------------------
// FAST-10. ref: Thirdparty/fast/src/fast_10.cpp:36-51,3164-3166 (scan bounds, strict compares),
// faster_corner_10_sse.cpp:27-32,182-196 (same set, raster order; small widths fall back to plain).
// ------------------------------------------------------------------------------------------
int orc_fast10_detect(const uint8_t* img, int w, int h, int stride, int barrier, int16_t* xy, int cap)
{
    int n = 0;
    // ref: faster_corner_10_sse.cpp:192-196 : width >= 22 but height < 7 returns nothing (same as empty scan)
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = img + (size_t)y * stride + x;
            const int cb = *p + barrier, c_b = *p - barrier;
            unsigned bright = 0, dark = 0;
            for (int k = 0; k < 16; ++k) {
                const int v = p[RING_DY[k] * stride + RING_DX[k]];
                if (v > cb) bright |= 1u << k;
                if (v < c_b) dark |= 1u << k;
            }
            if (has_arc10(bright) || has_arc10(dark)) {
                if (n < cap) { xy[2 * n] = (int16_t)x; xy[2 * n + 1] = (int16_t)y; }
                ++n;
            }
        }
    return n;
}

// ref: Thirdparty/fast/src/fast_10_score.cpp:21-31,3147 : largest barrier at which the pixel is still a corner
void orc_fast10_score(const uint8_t* img, int stride, const int16_t* xy, int n, int* scores)
{
    for (int i = 0; i < n; ++i) {
        const uint8_t* p = img + (size_t)xy[2 * i + 1] * stride + xy[2 * i];
        int d[16];
        for (int k = 0; k < 16; ++k) d[k] = (int)p[RING_DY[k] * stride + RING_DX[k]] - (int)*p;
        int best = INT_MIN;
        for (int s = 0; s < 16; ++s) {
            int mb = INT_MAX, md = INT_MAX;
            for (int k = 0; k < 10; ++k) {
                const int v = d[(s + k) & 15];
                mb = std::min(mb, v);
                md = std::min(md, -v);
            }
            best = std::max(best, std::max(mb, md));
        }
        scores[i] = best - 1;
    }
}

// ref: Thirdparty/fast/src/nonmax_3x3.cpp:17-112 : keep i iff no 8-neighbour corner has score >= own.
int orc_fast_nonmax_3x3(const int16_t* xy, const int* scores, int n, int* keep)
{
    // dense restatement: corners are in raster order, so a (y,x)->index map reproduces the list walk.
    if (n < 1) return 0;
    int maxx = 0, maxy = 0;
    for (int i = 0; i < n; ++i) { maxx = std::max(maxx, (int)xy[2 * i]); maxy = std::max(maxy, (int)xy[2 * i + 1]); }
    const int W = maxx + 3, H = maxy + 3;
    std::vector<int> map((size_t)W * H, -1);
    for (int i = 0; i < n; ++i) map[(size_t)(xy[2 * i + 1] + 1) * W + xy[2 * i] + 1] = i;
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const int x = xy[2 * i] + 1, y = xy[2 * i + 1] + 1;
        bool ok = true;
        for (int dy = -1; dy <= 1 && ok; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy) continue;
                const int j = map[(size_t)(y + dy) * W + x + dx];
                if (j >= 0 && scores[j] >= scores[i]) { ok = false; break; }
            }
        if (ok) keep[m++] = i;
    }
    return m;
}

// ------------------------------------------------------------------------------------------
// ref: src/Feature_detection.cpp:157-198
// ------------------------------------------------------------------------------------------
float orc_shitomasi(const uint8_t* img, int w, int h, int stride, int u, int v)
{
    float dXX = 0.0, dYY = 0.0, dXY = 0.0;
    const int halfbox_size = 4;
    const int box_size = 2 * halfbox_size;
    const int box_area = box_size * box_size;
    const int x_min = u - halfbox_size, x_max = u + halfbox_size;
    const int y_min = v - halfbox_size, y_max = v + halfbox_size;
    if (x_min < 1 || x_max >= w - 1 || y_min < 1 || y_max >= h - 1) return 0.0;
    for (int y = y_min; y < y_max; ++y) {
        const uint8_t* l = img + (size_t)stride * y + x_min - 1;
        const uint8_t* r = img + (size_t)stride * y + x_min + 1;
        const uint8_t* t = img + (size_t)stride * (y - 1) + x_min;
        const uint8_t* b = img + (size_t)stride * (y + 1) + x_min;
        for (int x = 0; x < box_size; ++x, ++l, ++r, ++t, ++b) {
            float dx = *r - *l;
            float dy = *b - *t;
            dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
        }
    }
    dXX = dXX / (2.0 * box_area);
    dYY = dYY / (2.0 * box_area);
    dXY = dXY / (2.0 * box_area);
    // std::sqrt(float) is the float overload in C++ (ref includes <math.h> via include/Camera.h:13, g++ >= 6)
    return 0.5 * (dXX + dYY - std::sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}

// ref: src/Feature_detection.cpp:69-109
void orc_detect_cells(const uint8_t* pyr, const int* offs, const int* ws, const int* hs, int levels,
                      int img_w, int img_h, int cell_size, const uint8_t* occupied, double thr, orc_corner* cells)
{
    const int grid_rows = (int)std::ceil(1.0 * img_h / cell_size);   // ref: :18-19
    const int grid_cols = (int)std::ceil(1.0 * img_w / cell_size);
    for (int i = 0; i < grid_rows * grid_cols; ++i) cells[i] = orc_corner{ 0, 0, 0, (float)thr };  // ref: :74
    std::vector<int16_t> xy;
    std::vector<int> scores, keep;
    for (int L = 0; L < levels; ++L) {
        const int scale = 1 << L;
        const uint8_t* img = pyr + offs[L];
        const int w = ws[L], h = hs[L];
        xy.resize((size_t)2 * w * h);
        const int n = orc_fast10_detect(img, w, h, w, 20, xy.data(), w * h);       // ref: :81-82 barrier literal 20
        scores.resize(n); keep.resize(n);
        orc_fast10_score(img, w, xy.data(), n, scores.data());                     // ref: :91
        const int m = orc_fast_nonmax_3x3(xy.data(), scores.data(), n, keep.data());  // ref: :92
        for (int q = 0; q < m; ++q) {
            const int x = xy[2 * keep[q]], y = xy[2 * keep[q] + 1];
            const int k = ((y * scale) / cell_size) * grid_cols + (x * scale) / cell_size;   // ref: :97-98
            if (occupied && occupied[k]) continue;                                     // ref: :100
            const float s = orc_shitomasi(img, w, h, w, x, y);                          // ref: :103
            if (s > cells[k].score) cells[k] = orc_corner{ x * scale, y * scale, L, s };  // ref: :104-107
        }
    }
}

// ------------------------------------------------------------------------------------------
// B.2 cv::circle filled, thickness -1, LINE_8, shift 0 -> OpenCV drawing.cpp Circle() midpoint routine.
// ------------------------------------------------------------------------------------------
static inline void hline(uint8_t* row, int x0, int x1, uint8_t c) { for (int x = x0; x <= x1; ++x) row[x] = c; }

void orc_circle_fill(uint8_t* img, int w, int h, int stride, int cx, int cy, int radius, uint8_t color)
{
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    const bool inside = cx >= radius && cx < w - radius && cy >= radius && cy < h - radius;
    while (dx >= dy) {
        const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        if (inside) {
            hline(img + (size_t)y11 * stride, x11, x12, color);
            hline(img + (size_t)y12 * stride, x11, x12, color);
            hline(img + (size_t)y21 * stride, x21, x22, color);
            hline(img + (size_t)y22 * stride, x21, x22, color);
        } else if (x11 < w && x12 >= 0 && y21 < h && y22 >= 0) {
            x11 = std::max(x11, 0);
            x12 = std::min(x12, w - 1);
            if ((unsigned)y11 < (unsigned)h) hline(img + (size_t)y11 * stride, x11, x12, color);
            if ((unsigned)y12 < (unsigned)h) hline(img + (size_t)y12 * stride, x11, x12, color);
            if (x21 < w && x22 >= 0) {
                x21 = std::max(x21, 0);
                x22 = std::min(x22, w - 1);
                if ((unsigned)y21 < (unsigned)h) hline(img + (size_t)y21 * stride, x21, x22, color);
                if ((unsigned)y22 < (unsigned)h) hline(img + (size_t)y22 * stride, x21, x22, color);
            }
        }
        dy++;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

// ref: src/Feature_detection.cpp:111-150
int orc_detect_select(orc_corner* cells, int n_cells, uint8_t* mask, int img_w, int img_h, int cell_size,
                      int max_fts, int n_existing, orc_corner* out)
{
    // ref: include/Feature_detection.h:29-32 operator< is "score descending"; std::sort is unstable on purpose (Q7)
    std::sort(cells, cells + n_cells, [](const orc_corner& a, const orc_corner& b) { return b.score < a.score; });
    int added = 0;
    for (int i = 0; i < n_cells; ++i) {
        const orc_corner c = cells[i];
        if (c.score > 20) {                                              // ref: :128 literal 20
            const int mx = orc_cvround((float)c.x), my = orc_cvround((float)c.y);   // Mat::at(Point2f)
            if (mask[(size_t)my * img_w + mx] == 255) {                  // ref: :142
                out[added++] = c;                                        // ref: :145
                orc_circle_fill(mask, img_w, img_h, img_w, mx, my, cell_size, 0);  // ref: :146
            }
        }
        if (n_existing + added >= max_fts) break;                        // ref: :148-149
    }
    return added;
}

// ------------------------------------------------------------------------------------------
// ref: src/Camera.cpp:173-178 (float evaluation), src/Frame.cpp:83-92 (normalize)
// ------------------------------------------------------------------------------------------
void orc_feature_normal(const orc_cam* cam, const float px[2], double normal[3])
{
    const float depth = 1.0f;
    double n[3] = { (double)(depth * (px[0] - cam->cx) / cam->fx), (double)(depth * (px[1] - cam->cy) / cam->fy), (double)depth };
    const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    normal[0] = n[0] / nn; normal[1] = n[1] / nn; normal[2] = n[2] / nn;
}

// ------------------------------------------------------------------------------------------
// ReprojectPoint / Get_ClosetObs / IsInImage
// ------------------------------------------------------------------------------------------
int orc_is_in_image(const orc_cam* cam, float x, float y, int boundary, int level)   // ref: src/Camera.cpp:187-193
{
    const int rx = orc_cvround(x), ry = orc_cvround(y);
    return rx >= boundary && rx < cam->width / (1 << level) - boundary && ry >= boundary && ry < cam->height / (1 << level) - boundary;
}

int orc_reproject_point(const orc_cam* cam, const double pose_cur_c2w[7], const double point_w[3], int cell_size, int grid_cols,
                        double px[2], int* cell)                                   // ref: src/Feature_alignment.cpp:54-69
{
    double q[3];
    se3_act(se3_from(pose_cur_c2w), point_w, q);                                   // ref: src/Frame.cpp:320
    px[0] = (double)cam->fx * q[0] / q[2] + (double)cam->cx;                       // ref: src/Camera.cpp:167-171
    px[1] = (double)cam->fy * q[1] / q[2] + (double)cam->cy;
    if (!orc_is_in_image(cam, (float)px[0], (float)px[1], 8, 0)) return 0;         // ref: :58
    *cell = static_cast<int>(px[1] / cell_size) * grid_cols + static_cast<int>(px[0] / cell_size);   // ref: :60-61
    return 1;
}

int orc_closest_obs(const double cur_center[3], const double point_w[3], const double* kf_centers, int n_obs, int* best)
{                                                                                  // ref: src/MapPoint.cpp:133-174
    if (n_obs <= 0) { *best = -1; return 0; }
    double f[3] = { cur_center[0] - point_w[0], cur_center[1] - point_w[1], cur_center[2] - point_w[2] };
    double n = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    f[0] /= n; f[1] /= n; f[2] /= n;                                               // Eigen normalize(): *this /= norm()
    double min_angle = 0;
    int it = 0;
    for (int j = 0; j < n_obs; ++j) {
        double r[3] = { kf_centers[3 * j] - point_w[0], kf_centers[3 * j + 1] - point_w[1], kf_centers[3 * j + 2] - point_w[2] };
        const double rn = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        r[0] /= rn; r[1] /= rn; r[2] /= rn;
        const double c = r[0] * f[0] + r[1] * f[1] + r[2] * f[2];
        if (c > min_angle) { min_angle = c; it = j; }                              // ref: :154-158
    }
    *best = it;
    return !(min_angle < 0.5);                                                     // ref: :170-171
}

// ------------------------------------------------------------------------------------------
// keyframe ingest
// ------------------------------------------------------------------------------------------
void orc_undistort_points(const orc_cam* cam, const float dist[5], const float* src, int n, float* dst)
{
    // OpenCV cvUndistortPointsInternal (imgproc/src/undistort.dispatch.cpp), CV_32F K / dist widened to double, no R,
    // P = K, criteria = (MAX_ITER, 5): not under /root/reference, restated from the published algorithm.
    const double fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k0 = dist[0], k1 = dist[1], p1 = dist[2], p2 = dist[3], k2 = dist[4];   // k[0] k[1] k[2] k[3] k[4]; k[5..] = 0
    for (int i = 0; i < n; ++i) {
        double x = src[2 * i], y = src[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;                                               // identity tilt: invProj = 1
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k2 * r2 + k1) * r2 + k0) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0 * r2 + 0 * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        // RR = P * I = K : xx = fx*x + 0*y + cx, ww = 1/(0*x + 0*y + 1)
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        dst[2 * i] = (float)(xx * ww);
        dst[2 * i + 1] = (float)(yy * ww);
    }
}

void orc_depth_convert(const uint16_t* depth, int n, float depth_scale, float* out)  // ref: src/Tracking.cpp:56
{
    const float a = (float)(double)(1.0f / depth_scale);                           // alpha travels as double, cvt uses float
    for (int i = 0; i < n; ++i) out[i] = (float)depth[i] * a;
}

float orc_feature_depth(const float* depth, int w, int h, const float px[2])       // ref: src/Frame.cpp:200-224
{
    const int x = orc_cvround(px[0]), y = orc_cvround(px[1]);
    auto at = [&](int xx, int yy) -> float { return (xx < 0 || yy < 0 || xx >= w || yy >= h) ? 0.f : depth[(size_t)yy * w + xx]; };
    float d = at(x, y);
    if (d != 0) return d;
    const int dx[4] = { -1, 0, 1, 0 }, dy[4] = { 0, -1, 0, 1 };
    for (int i = 0; i < 4; ++i) {
        d = at(x + dx[i], y + dy[i]);
        if (d != 0) return d;
    }
    return -1.0f;
}

void orc_unproject(const orc_cam* cam, const double pose_c2w[7], const float px[2], float d, double out[3])
{                                                                                  // ref: src/Frame.cpp:152-157
    const double p[3] = { (double)(d * (px[0] - cam->cx) / cam->fx), (double)(d * (px[1] - cam->cy) / cam->fy), (double)d };
    se3_act(se3_inv(se3_from(pose_c2w)), p, out);
}

// ------------------------------------------------------------------------------------------
// cv::CLAHE::apply, CV_8UC1 (OpenCV imgproc/src/clahe.cpp: CLAHE_CalcLut_Body + CLAHE_Interpolation_Body), sizes divisible
// by the tile grid. Integer histogram / clip / redistribution, float LUT scale and float bilinear blend of four tile LUTs.
// ------------------------------------------------------------------------------------------
void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst)
{
    const int tw = w / tiles_x, th = h / tiles_y, total = tw * th, hist_size = 256;
    const float lut_scale = static_cast<float>(hist_size - 1) / total;
    int clip = 0;
    if (clip_limit > 0.0) {
        clip = static_cast<int>(clip_limit * total / hist_size);
        clip = std::max(clip, 1);
    }
    std::vector<uint8_t> lut((size_t)tiles_x * tiles_y * hist_size);
    for (int ty = 0; ty < tiles_y; ++ty)
        for (int tx = 0; tx < tiles_x; ++tx) {
            int hist[256] = { 0 };
            for (int y = 0; y < th; ++y) {
                const uint8_t* row = src + (size_t)(ty * th + y) * w + tx * tw;
                for (int x = 0; x < tw; ++x) hist[row[x]]++;
            }
            if (clip > 0) {
                int clipped = 0;
                for (int i = 0; i < hist_size; ++i)
                    if (hist[i] > clip) { clipped += hist[i] - clip; hist[i] = clip; }
                const int batch = clipped / hist_size;
                int residual = clipped - batch * hist_size;
                for (int i = 0; i < hist_size; ++i) hist[i] += batch;
                if (residual != 0) {
                    const int step = std::max(hist_size / residual, 1);
                    for (int i = 0; i < hist_size && residual > 0; i += step, residual--) hist[i]++;
                }
            }
            uint8_t* l = &lut[(size_t)(ty * tiles_x + tx) * hist_size];
            int sum = 0;
            for (int i = 0; i < hist_size; ++i) {
                sum += hist[i];
                const int v = orc_cvround((float)sum * lut_scale);      // saturate_cast<uchar>(float): cvRound + clamp
                l[i] = (uint8_t)std::min(std::max(v, 0), 255);
            }
        }
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < h; ++y) {
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int)std::floor(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        ty1 = std::max(ty1, 0); ty2 = std::min(ty2, tiles_y - 1);
        const uint8_t* p1 = &lut[(size_t)ty1 * tiles_x * hist_size];
        const uint8_t* p2 = &lut[(size_t)ty2 * tiles_x * hist_size];
        for (int x = 0; x < w; ++x) {
            const float txf = x * inv_tw - 0.5f;
            int tx1 = (int)std::floor(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            tx1 = std::max(tx1, 0); tx2 = std::min(tx2, tiles_x - 1);
            const int v = src[(size_t)y * w + x];
            const int i1 = tx1 * hist_size + v, i2 = tx2 * hist_size + v;
            const float res = (p1[i1] * xa1 + p1[i2] * xa) * ya1 + (p2[i1] * xa1 + p2[i2] * xa) * ya;
            const int r = orc_cvround(res);
            dst[(size_t)y * w + x] = (uint8_t)std::min(std::max(r, 0), 255);
        }
    }
}

void orc_ldlt6_solve(const double H[36], const double b[6], double x[6]) { ldlt6_solve(H, b, x); }
void orc_se3_exp(const double x[6], double pose[7]) { se3_to(se3_exp(x), pose); }
void orc_se3_mul(const double a[7], const double b[7], double out[7]) { se3_to(se3_mul(se3_from(a), se3_from(b)), out); }
void orc_se3_inv(const double a[7], double out[7]) { se3_to(se3_inv(se3_from(a)), out); }
void orc_se3_act(const double a[7], const double p[3], double out[3]) { se3_act(se3_from(a), p, out); }

// ------------------------------------------------------------------------------------------
// Sparse image alignment. ref: src/Sprase_ImageAlign.cpp
// ------------------------------------------------------------------------------------------
namespace {
struct SparseAlign {
    const orc_cam* cam;
    const uint8_t *ref_pyr, *cur_pyr;
    const int *offs, *ws, *hs;
    const orc_ref_feat* feats; int n_feats;
    double ref_center[3];
    static const int kHalf = 4;                 // ref: include/Feature_alignment.h:21
    static const int kArea = kHalf * kHalf;     // ref: src/Sprase_ImageAlign.cpp:12 mPatchArea
    std::vector<double> ref_patch;              // mRefPatch  N x 16 row-major   (ref: include/Sprase_ImageAlign.h:61)
    std::vector<double> jac;                    // mJocabianPatch (N*16) x 6     (ref: :62)
    std::vector<double> ref_pts;                // mRefNormals 3 x N, column = point (ref: :63)
    int n_ref = 0;
    double H[36], JRes[6];

    static void jacobian_ba(const double p[3], double J[12])  // ref: :169-193
    {
        const double x = p[0], y = p[1];
        const double z_inv = 1.0 / p[2];
        const double z_inv2 = z_inv * z_inv;
        J[0] = -z_inv; J[1] = 0.0; J[2] = x * z_inv2; J[3] = y * J[2]; J[4] = -(1.0 + x * J[2]); J[5] = y * z_inv;
        J[6] = 0.0; J[7] = -z_inv; J[8] = y * z_inv2; J[9] = 1.0 + y * J[8]; J[10] = -x * J[8]; J[11] = -x * z_inv;
    }

    void precompute(int level)  // GetJocabianMat, ref: :62-166
    {
        const uint8_t* img = ref_pyr + offs[level];
        const int cols = ws[level], rows = hs[level], step = ws[level];
        const float tScale = 1.0 / (1 << level);
        const int boarder = 0.5 * kHalf + 1;
        const float tFocalth = cam->f;          // Q1: Camera.f, not fx/fy
        std::vector<double> pts, normals, points;
        n_ref = 0;
        for (int i = 0; i < n_feats; ++i) {
            const orc_ref_feat& f = feats[i];
            if (!f.initial) continue;                                               // ref: :86
            const double px = (double)f.px[0] * tScale, py = (double)f.px[1] * tScale;   // ref: :89-91
            const bool zero = (f.point_w[0] == 0 && f.point_w[1] == 0 && f.point_w[2] == 0);  // isZero(0)
            if (zero || px - boarder < 0 || py - boarder < 0 || px + boarder >= cols || py + boarder >= rows) continue;  // ref: :95-100
            pts.push_back(px); pts.push_back(py);
            for (int k = 0; k < 3; ++k) { points.push_back(f.point_w[k]); normals.push_back(f.normal[k]); }
            ++n_ref;
        }
        ref_patch.assign((size_t)n_ref * kArea, 0.0);
        jac.assign((size_t)n_ref * kArea * 6, 0.0);
        ref_pts.assign((size_t)n_ref * 3, 0.0);
        for (int j = 0; j < n_ref; ++j) {
            // ref: :117-119  P_ref = mNormal * ||P_w - O_ref||
            const double d0 = points[3 * j] - ref_center[0], d1 = points[3 * j + 1] - ref_center[1], d2 = points[3 * j + 2] - ref_center[2];
            const double depth = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            for (int k = 0; k < 3; ++k) ref_pts[3 * j + k] = normals[3 * j + k] * depth;
            // ref: :123-132
            const int fx_ = (int)std::floor(pts[2 * j]), fy_ = (int)std::floor(pts[2 * j + 1]);
            const double sx = pts[2 * j] - fx_, sy = pts[2 * j + 1] - fy_;
            const double w00 = (1.0 - sx) * (1.0 - sy), w01 = sx * (1.0 - sy), w10 = (1.0 - sx) * sy, w11 = sx * sy;
            double J[12];
            jacobian_ba(&ref_pts[3 * j], J);
            int num = 0;
            for (int r = 0; r < kHalf; ++r) {
                const uint8_t* it = img + (size_t)(fy_ - 2 + r) * step + (fx_ - 2);
                for (int c = 0; c < kHalf; ++c, ++it, ++num) {
                    ref_patch[(size_t)j * kArea + num] = w00 * it[0] + w01 * it[1] + w10 * it[step] + w11 * it[step + 1];   // ref: :147-148
                    const double dx = 0.5 * ((w00 * it[1] + w01 * it[2] + w10 * it[step + 1] + w11 * it[step + 2]) -
                                             (w00 * it[-1] + w01 * it[0] + w10 * it[step - 1] + w11 * it[step]));          // ref: :150-153
                    const double dy = 0.5 * ((w00 * it[step] + w01 * it[step + 1] + w10 * it[2 * step] + w11 * it[2 * step + 1]) -
                                             (w00 * it[-step] + w01 * it[-step + 1] + w10 * it[0] + w11 * it[1]));          // ref: :155-158
                    double* row = &jac[((size_t)j * kArea + num) * 6];
                    for (int k = 0; k < 6; ++k) row[k] = (dx * J[k] + dy * J[6 + k]) * (double)tFocalth * (double)tScale;     // ref: :160
                }
            }
        }
    }

    double residuals(const Se3& T, int level, bool linear, int& n_pts)  // ComputeResiduals, ref: :240-299
    {
        const uint8_t* img = cur_pyr + offs[level];
        const int cols = ws[level], rows = hs[level];
        const float tScale = 1.0 / (1 << level);
        const int border = kHalf - 1;
        double chi2 = 0.0;
        int res_num = 0;
        n_pts = 0;
        for (int n = 0; n < n_ref; ++n) {
            double q[3];
            se3_act(T, &ref_pts[3 * n], q);                                                   // ref: :254
            // Camera2Pixel (ref: src/Camera.cpp:167-171) with float intrinsics widened, then * tScale
            const double u = ((double)cam->fx * q[0] / q[2] + (double)cam->cx) * (double)tScale;
            const double v = ((double)cam->fy * q[1] / q[2] + (double)cam->cy) * (double)tScale;
            const int u_i = (int)std::floor(u), v_i = (int)std::floor(v);
            if (u_i < 0 || v_i < 0 || u_i - border < 0 || v_i - border < 0 || u_i + border >= cols || v_i + border >= rows) continue;  // ref: :262
            const double su = u - u_i, sv = v - v_i;
            const double tl = (1.0 - su) * (1.0 - sv), tr = su * (1.0 - sv), bl = (1.0 - su) * sv, br = su * sv;
            const int step = cols;   // Q5: continuous level images
            int num = 0;
            for (int i = 0; i < kHalf; ++i) {
                const uint8_t* it = img + (size_t)(v_i + i - 2) * cols + u_i - 2;
                for (int j = 0; j < kHalf; ++j, ++it, ++num) {
                    const double cur = tl * it[0] + tr * it[1] + bl * it[step] + br * it[step + 1];   // ref: :281
                    const double res = -(ref_patch[(size_t)n * kArea + num] - cur);                  // ref: :282
                    chi2 += res * res;
                    res_num++;
                    if (linear) {
                        const double* J = &jac[((size_t)n * kArea + num) * 6];
                        for (int a = 0; a < 6; ++a) {
                            for (int b = 0; b < 6; ++b) H[a * 6 + b] += J[a] * J[b];                  // ref: :290
                            JRes[a] += J[a] * res;                                                    // ref: :291
                        }
                    }
                }
            }
            n_pts++;
        }
        return chi2 / res_num;   // Q2: NaN when nothing is visible
    }
};
}  // namespace

int orc_sparse_align(const orc_cam* cam, const uint8_t* ref_pyr, const uint8_t* cur_pyr, const int* offs, const int* ws,
                     const int* hs, const orc_ref_feat* feats, int n_feats, const double ref_center[3],
                     const double pose_in[7], int max_level, int min_level, int max_iters, double pose_out[7],
                     orc_iter_log* log, int log_cap, int* n_log)
{
    SparseAlign sa;
    sa.cam = cam; sa.ref_pyr = ref_pyr; sa.cur_pyr = cur_pyr; sa.offs = offs; sa.ws = ws; sa.hs = hs;
    sa.feats = feats; sa.n_feats = n_feats;
    for (int k = 0; k < 3; ++k) sa.ref_center[k] = ref_center[k];
    Se3 T = se3_from(pose_in);
    int n_pts = 0, nl = 0;
    for (int level = max_level - 1; level >= min_level; --level) {        // ref: :45
        sa.precompute(level);                                             // ref: :47
        // GaussNewtonSolver, ref: :301-344
        bool stop = false;
        const double eps = 1e-8;
        double chi2 = 0.0;
        Se3 Told = T;
        for (int i = 0; i < max_iters; ++i) {
            for (int k = 0; k < 36; ++k) sa.H[k] = 0;
            for (int k = 0; k < 6; ++k) sa.JRes[k] = 0;
            const double chi2New = sa.residuals(T, level, true, n_pts);   // ref: :317
            double x[6];
            ldlt6_solve(sa.H, sa.JRes, x);                                 // ref: :318
            int flags = 0;
            if (std::isnan(x[0])) { stop = true; flags |= 4; }             // ref: :321-326
            orc_iter_log* e = (log && nl < log_cap) ? &log[nl] : nullptr;
            if (e) { e->level = level; e->iter = i; e->n_pts = n_pts; e->chi2 = chi2New; for (int k = 0; k < 6; ++k) e->x[k] = x[k]; }
            if ((i > 0 && chi2New > chi2) || stop) {                       // ref: :328-332
                T = Told;
                if (e) e->flags = flags | 2;
                ++nl;
                break;
            }
            Se3 Tnew = se3_mul(T, se3_exp(x));                             // ref: :335
            Told = T;
            T = Tnew;
            chi2 = chi2New;
            flags |= 1;
            double mx = 0; for (int k = 0; k < 6; ++k) mx = std::max(mx, std::fabs(x[k]));
            const bool conv = mx <= eps;                                   // ref: :341
            if (conv) flags |= 8;
            if (e) e->flags = flags;
            ++nl;
            if (conv) break;
        }
    }
    se3_to(T, pose_out);
    if (n_log) *n_log = nl;
    return n_pts;
}

// ------------------------------------------------------------------------------------------
// Feature alignment. ref: src/Feature_alignment.cpp
// ------------------------------------------------------------------------------------------
static inline void cam2pix(const orc_cam* cam, const double p[3], double px[2])  // ref: src/Camera.cpp:167-171
{
    px[0] = (double)cam->fx * p[0] / p[2] + (double)cam->cx;
    px[1] = (double)cam->fy * p[1] / p[2] + (double)cam->cy;
}
static inline void pix2cam_d(const orc_cam* cam, const double px[2], float depth, double out[3])  // ref: src/Camera.cpp:180-185
{
    out[0] = depth * (px[0] - cam->cx) / cam->fx;
    out[1] = depth * (px[1] - cam->cy) / cam->fy;
    out[2] = depth;
}

void orc_solve_affine(const orc_cam* cam, const double kf_center[3], const double ref_point_w[3], const double ref_normal[3],
                      const float ref_px[2], int ref_level, const double pose_c2r[7], double A[4])  // ref: :160-190
{
    const int HalfLarger = 4 + 1;
    const double d0 = kf_center[0] - ref_point_w[0], d1 = kf_center[1] - ref_point_w[1], d2 = kf_center[2] - ref_point_w[2];
    const double nrm = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double P[3] = { nrm * ref_normal[0], nrm * ref_normal[1], nrm * ref_normal[2] };   // ref: :167
    // ref: :171-172  (float px + int) evaluated in float, then widened
    const double pxU[2] = { (double)(ref_px[0] + HalfLarger * (1 << ref_level)), (double)ref_px[1] };
    const double pxV[2] = { (double)ref_px[0], (double)(ref_px[1] + HalfLarger * (1 << ref_level)) };
    double PU[3], PV[3];
    pix2cam_d(cam, pxU, 1.0f, PU);
    pix2cam_d(cam, pxV, 1.0f, PV);
    double n = std::sqrt(PU[0] * PU[0] + PU[1] * PU[1] + PU[2] * PU[2]); PU[0] /= n; PU[1] /= n; PU[2] /= n;
    n = std::sqrt(PV[0] * PV[0] + PV[1] * PV[1] + PV[2] * PV[2]);        PV[0] /= n; PV[1] /= n; PV[2] /= n;
    double s = P[2] / PU[2]; PU[0] *= s; PU[1] *= s; PU[2] *= s;          // ref: :178
    s = P[2] / PV[2];        PV[0] *= s; PV[1] *= s; PV[2] *= s;          // ref: :179
    const Se3 T = se3_from(pose_c2r);
    double q[3], c[2], cu[2], cv[2];
    se3_act(T, P, q);  cam2pix(cam, q, c);
    se3_act(T, PU, q); cam2pix(cam, q, cu);
    se3_act(T, PV, q); cam2pix(cam, q, cv);
    A[0] = (cu[0] - c[0]) / HalfLarger; A[2] = (cu[1] - c[1]) / HalfLarger;   // col 0
    A[1] = (cv[0] - c[0]) / HalfLarger; A[3] = (cv[1] - c[1]) / HalfLarger;   // col 1
}

int orc_best_search_level(const double A[4], int max_level)  // ref: :192-204
{
    int L = 0;
    double D = A[0] * A[3] - A[2] * A[1];   // Eigen 2x2 determinant: m00*m11 - m10*m01
    while (D > 3.0 && L < max_level) { L++; D = D * 0.25; }
    return L;
}

void orc_warp_affine(const double A[4], const uint8_t* img, int w, int h, int stride, const float ref_px_in[2], int ref_level,
                     int search_level, uint8_t patch[100])  // ref: :206-259
{
    // Eigen 2x2 inverse (double), then cast<float>
    const double det = A[0] * A[3] - A[2] * A[1];
    const double invdet = 1.0 / det;
    const float a00 = (float)(A[3] * invdet), a01 = (float)(-A[1] * invdet), a10 = (float)(-A[2] * invdet), a11 = (float)(A[0] * invdet);
    const float rx = ref_px_in[0] / (1 << ref_level), ry = ref_px_in[1] / (1 << ref_level);   // ref: :215-216
    const float k = (float)(1 / (1 << search_level));                                           // Q3: integer division
    for (int j = 0; j < 100; ++j) {
        const float gx = (float)(j % 10 - 5), gy = (float)(j / 10 - 5);                          // ref: :220-228
        float wx = (a00 * gx + a01 * gy) * k;                                                    // ref: :231
        float wy = (a10 * gx + a11 * gy) * k;
        wx = wx + rx; wy = wy + ry;                                                              // ref: :232
        const int fx_ = (int)std::floor((double)wx), fy_ = (int)std::floor((double)wy);          // Eigenfloor(double)
        const float sx = wx - (float)fx_, sy = wy - (float)fy_;
        const float ox = 1.0f - sx, oy = 1.0f - sy;
        const float W00 = ox * oy, W01 = ox * sy, W10 = sx * oy;
        const float W11 = 1.0f - W00 - W01 - W10;                                                // ref: :244
        if (wx < 0 || wy < 0 || wx > w - 1 || wy > h - 1) {                                      // ref: :249
            patch[j] = 0;
        } else {
            // memory-safety note: with wx == w-1 (or wy == h-1) the reference reads one element past the row/image
            // with weight exactly 0; we read 0 instead (same result for any finite neighbour).
            auto at = [&](int x, int y) -> int { return (x < w && y < h) ? img[(size_t)y * stride + x] : 0; };
            const float v = W00 * at(fx_, fy_) + W01 * at(fx_, fy_ + 1) + W10 * at(fx_ + 1, fy_) + W11 * at(fx_ + 1, fy_ + 1);   // ref: :254-255
            patch[j] = (uint8_t)v;   // truncating store
        }
    }
}

void orc_patch_no_border(const uint8_t p10[100], uint8_t p8[64])  // ref: :261-275
{
    for (int i = 1; i < 9; ++i)
        for (int j = 0; j < 8; ++j) p8[(i - 1) * 8 + j] = p10[i * 10 + 1 + j];
}

int orc_align2d(const uint8_t* img, int w, int h, int stride, const uint8_t p10[100], const uint8_t p8[64], int max_iters,
                double px[2], int* n_iters_out)  // ref: :318-417
{
    const int PS = 8, LPS = 10, half = 4;
    float H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    float rdx[64], rdy[64];
    int idx = 0;
    for (int l = 0; l < PS; ++l) {
        const uint8_t* it = p10 + (l + 1) * LPS + 1;
        for (int i = 0; i < PS; ++i, ++it, ++idx) {
            float J[3];
            J[0] = 0.5 * (it[1] - it[-1]);
            J[1] = 0.5 * (it[LPS] - it[-LPS]);
            J[2] = 1;
            rdx[idx] = J[0]; rdy[idx] = J[1];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[a * 3 + b] += J[a] * J[b];   // ref: :341
        }
    }
    // Eigen 3x3 inverse: cofactors / determinant (App. B.4)
    float Hinv[9];
    {
        auto m = [&](int r, int c) { return H[r * 3 + c]; };
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
        };
        const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const float det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
        const float invdet = 1.0f / det;
        Hinv[0] = c00 * invdet; Hinv[1] = c10 * invdet; Hinv[2] = c20 * invdet;
        Hinv[3] = cof(0, 1) * invdet; Hinv[4] = cof(1, 1) * invdet; Hinv[5] = cof(2, 1) * invdet;
        Hinv[6] = cof(0, 2) * invdet; Hinv[7] = cof(1, 2) * invdet; Hinv[8] = cof(2, 2) * invdet;
    }
    float mean_diff = 0;
    float u = px[0], v = px[1];
    const float min_update_squared = 0.03 * 0.03;
    bool converged = false;
    int it_count = 0;
    const size_t img_bytes = (size_t)stride * h;
    for (int i = 0; i < max_iters; ++i) {
        const int u_r = (int)std::floor(u), v_r = (int)std::floor(v);
        if (u_r < half || v_r < half || u_r > w - half || v_r > h - half || std::isnan(u) || std::isnan(v)) break;   // ref: :367-369 (Q4: '>' not '>=')
        ++it_count;
        const float sx = u - u_r, sy = v - v_r;
        const float wTL = (1.0 - sx) * (1.0 - sy);
        const float wTR = sx * (1 - sy);
        const float wBL = (1.0 - sx) * sy;
        const float wBR = sx * sy;
        float Jres[3] = { 0, 0, 0 };
        int k2 = 0;
        for (int j = 0; j < PS; ++j) {
            const size_t base = (size_t)(v_r + j - half) * stride + u_r - half;
            for (int k = 0; k < PS; ++k, ++k2) {
                // Q4 memory-safety rule: linear addressing; bytes past the end of the level image read as 0.
                auto at = [&](size_t o) -> int { return o < img_bytes ? img[o] : 0; };
                const size_t o = base + k;
                const float s = wTL * at(o) + wTR * at(o + 1) + wBL * at(o + stride) + wBR * at(o + stride + 1);   // ref: :386
                const float res = s - p8[k2] + mean_diff;                                                           // ref: :387
                Jres[0] -= res * rdx[k2];
                Jres[1] -= res * rdy[k2];
                Jres[2] -= res;
            }
        }
        float upd[3];
        for (int a = 0; a < 3; ++a) upd[a] = Hinv[a * 3] * Jres[0] + Hinv[a * 3 + 1] * Jres[1] + Hinv[a * 3 + 2] * Jres[2];   // ref: :395
        u += upd[0]; v += upd[1]; mean_diff += upd[2];
        if (upd[0] * upd[0] + upd[1] * upd[1] < min_update_squared) { converged = true; break; }   // ref: :400
    }
    px[0] = u; px[1] = v;
    if (n_iters_out) *n_iters_out = it_count;
    return converged ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// SURVEY 8f-1, the caller's half: Tracking::GetCloseKeyFrames (ref: src/Tracking.cpp:315-345) and the ranking at the top of
// Tracking::UpdateLocalMap (ref: :261-277). A key frame is "close" when ANY of its map points is visible in the current frame
// (Frame::isVisible, ref: src/Frame.cpp:300-311: in front of the camera and Camera::IsInImage of the float pixel, boundary 0);
// null and exactly-zero points are skipped (:324-328); its rank key is the distance between the two poses' TRANSLATIONS (:332,
// not the camera centres). UpdateLocalMap sorts the list (std::list::sort: stable) and keeps the first max_local (10).
// kfs are visited in the caller's order (std::set<KeyFrame*> iteration order in the reference). PARITY UNPINNED (no golden).
// ------------------------------------------------------------------------------------------
extern "C" int orc_close_keyframes(const orc_cam* cam, const double pose_cur_c2w[7], const int* pt_begin, const int* pt_count,
                                   const double* kf_t, int n_kfs, const double* points_w, int max_local, uint8_t* visible,
                                   double* dist, int* local)
{
    const Se3 T = se3_from(pose_cur_c2w);
    std::vector<std::pair<int, double>> close;
    for (int k = 0; k < n_kfs; ++k) {
        visible[k] = 0; dist[k] = 0.0;
        for (int i = 0; i < pt_count[k]; ++i) {
            const double* P = points_w + 3 * (size_t)(pt_begin[k] + i);
            if (P[0] == 0.0 && P[1] == 0.0 && P[2] == 0.0) continue;          // null map point or isZero(0)
            double c[3];
            se3_act(T, P, c);
            if (c[2] < 0.0) continue;                                           // ref: src/Frame.cpp:303
            double px[2];
            cam2pix(cam, c, px);
            if (!orc_is_in_image(cam, (float)px[0], (float)px[1], 0, 0)) continue;
            const double d0 = T.t[0] - kf_t[3 * k], d1 = T.t[1] - kf_t[3 * k + 1], d2 = T.t[2] - kf_t[3 * k + 2];
            visible[k] = 1;
            dist[k] = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            close.emplace_back(k, dist[k]);
            break;
        }
    }
    std::stable_sort(close.begin(), close.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.second < b.second; });
    int n = 0;
    for (const auto& c : close) { if (n >= max_local) break; local[n++] = c.first; }
    return n;
}

// ------------------------------------------------------------------------------------------
// Optimizer::PoseOptimization. ref: src/Optimizer.cpp:20-101, include/Optimizer.h:129-258.
//
// The reference hands the problem to ceres::Solve (ceres-solver is a find_package dependency, CMakeLists.txt:23, version
// unpinned, not under /root/reference and not installed here) with: one 6-vector parameter block [t, log(R)] carrying
// PoseLocalParameterization (Plus = left multiplication by SE3(SO3::exp(d[3..5]), d[0..2]), ComputeJacobian = identity),
// one constant 3-vector block per map point, FullBA_Problem residuals (2 rows) under CauchyLoss(1.0), DENSE_SCHUR,
// max_num_iterations = 100 (the tIterations argument is ignored), everything else at Solver::Options defaults.
// What is restated below is ceres-solver's published trust-region algorithm for exactly that configuration
// (1.10 .. 1.14 agree on it; file names of 1.13): trust_region_minimizer.cc (IterationZero, the main loop, the three
// tolerances, Jacobi scaling 1/(1+sqrt(colnorm2)) fixed at iteration 0), levenberg_marquardt_strategy.cc (D = sqrt(clamp(
// diag(J'J),1e-6,1e32)/radius), radius /= max(1/3, 1-(2q-1)^3) on success, radius /= decrease_factor (2,4,8..) on
// failure), residual_block.cc + corrector.cc (cost = rho(s)/2; rho'' <= 0 for Cauchy => residual and Jacobian are both
// scaled by sqrt(rho')), loss_function.cc (CauchyLoss), schur_eliminator_impl.h with one e-block and no f-block
// (step = LLT-inverse(D^2 + sum e'e) * sum e'b), program_evaluator.h (sequential sums, num_threads = 1).
// PARITY UNPINNED: no golden value exists in the reference (Test/test_Optimizer.cpp prints nothing checkable).
// Self-consistency checks live in tests/test_oracle_golden.py (scipy least_squares(loss='cauchy') reaches the same optimum).
// ------------------------------------------------------------------------------------------
namespace {

const double SOPHUS_SMALL_EPS = 1e-10;

Quat so3_exp(const double om[3])  // Sophus SO3::expAndTheta (so3.cpp), then the normalising SO3(Quaternion) constructor
{
    const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double half = 0.5 * theta;
    double imag;
    const double real = std::cos(half);
    if (theta < SOPHUS_SMALL_EPS) {
        const double t2 = theta * theta, t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = std::sin(half) / theta;
    }
    Quat q = { real, imag * om[0], imag * om[1], imag * om[2] };
    qnormalize(q);
    return q;
}

void so3_log(const Quat& q, double out[3])  // Sophus SO3::logAndTheta (atan-based; the |w| < eps branch is overwritten there too)
{
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    const double w = q.w;
    const double squared_w = w * w;
    double f;
    if (n < SOPHUS_SMALL_EPS) f = 2. / w - 2. * (n * n) / (w * squared_w);
    else f = 2 * std::atan(n / w) / n;
    out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

struct PoseBA {
    int n;
    const double* normals;   // Feature::mNormal
    const int* levels;       // Feature::mlevel
    const double* points;    // MapPoint::Get_Pose()
    std::vector<double> J;   // n x 2 x 6 row-major, loss-corrected (and column-scaled once scaled() ran)
    std::vector<double> r;   // n x 2, loss-corrected
    double grad[6];

    // FullBA_Problem::Evaluate (ref: include/Optimizer.h:141-205) on residual block k; raw residual and, if wanted, raw Jacobian
    void block(const Quat& q, const double t[3], int k, double res[2], double* jac) const
    {
        double c[3];
        qrot(q, points + 3 * k, c);
        c[0] += t[0]; c[1] += t[1]; c[2] += t[2];
        const double pred[2] = { c[0] / c[2], c[1] / c[2] };
        const double* nm = normals + 3 * k;
        const double obs[2] = { nm[0] / nm[2], nm[1] / nm[2] };
        const double div = (double)(1 << levels[k]);
        res[0] = (obs[0] - pred[0]) / div;
        res[1] = (obs[1] - pred[1]) / div;
        if (jac) {
            const double x = c[0], y = c[1];
            const double z_inv = 1.0 / c[2];
            const double z_inv2 = z_inv * z_inv;
            jac[0] = -z_inv; jac[1] = 0.0; jac[2] = x * z_inv2; jac[3] = y * jac[2]; jac[4] = -(1.0 + x * jac[2]); jac[5] = y * z_inv;
            jac[6] = 0.0; jac[7] = -z_inv; jac[8] = y * z_inv2; jac[9] = 1.0 + y * jac[8]; jac[10] = -x * jac[8]; jac[11] = -x * z_inv;
        }
    }
    // ProgramEvaluator::Evaluate: cost (and, with want_jac, corrected residuals / Jacobian / gradient). false on non-finite values.
    bool evaluate(const double x[6], bool want_jac, double* cost)
    {
        const Quat q = so3_exp(x + 3);
        double c = 0.0;
        if (want_jac) for (int i = 0; i < 6; ++i) grad[i] = 0.0;
        for (int k = 0; k < n; ++k) {
            double res[2];
            double* jac = want_jac ? &J[12 * (size_t)k] : nullptr;
            block(q, x, k, res, jac);
            const double s = res[0] * res[0] + res[1] * res[1];
            // CauchyLoss(1.0): b = 1, c = 1
            const double sum = 1.0 + s * 1.0;
            const double inv = 1.0 / sum;
            const double rho0 = 1.0 * std::log(sum);
            const double rho1 = std::max(std::numeric_limits<double>::min(), inv);
            c += 0.5 * rho0;
            if (!std::isfinite(rho0)) { *cost = c; return false; }
            if (want_jac) {
                const double sqrt_rho1 = std::sqrt(rho1);   // Corrector: rho[2] = -inv*inv <= 0 => alpha = 0, scaling = sqrt(rho')
                for (int i = 0; i < 12; ++i) jac[i] *= sqrt_rho1;
                r[2 * k] = res[0] * sqrt_rho1; r[2 * k + 1] = res[1] * sqrt_rho1;
                for (int col = 0; col < 6; ++col) {
                    double tmp = 0.0;
                    tmp += jac[col] * r[2 * k];
                    tmp += jac[6 + col] * r[2 * k + 1];
                    grad[col] += tmp;
                }
            }
        }
        *cost = c;
        return std::isfinite(c);
    }
};

// PoseLocalParameterization::Plus (ref: include/Optimizer.h:220-236)
void pose_plus(const double x[6], const double d[6], double out[6])
{
    Se3 told, tdelta;
    told.q = so3_exp(x + 3); told.t[0] = x[0]; told.t[1] = x[1]; told.t[2] = x[2];
    tdelta.q = so3_exp(d + 3); tdelta.t[0] = d[0]; tdelta.t[1] = d[1]; tdelta.t[2] = d[2];
    const Se3 tnew = se3_mul(tdelta, told);
    out[0] = tnew.t[0]; out[1] = tnew.t[1]; out[2] = tnew.t[2];
    so3_log(tnew.q, out + 3);
}

// Eigen LLT (lower, unblocked) of a 6x6 SPD matrix, then M^-1 = llt.solve(I) and y = M^-1 g (InvertPSDMatrix + product)
bool llt6_inverse_times(const double M[36], const double g[6], double y[6])
{
    double L[36];
    std::memcpy(L, M, sizeof(L));
    for (int k = 0; k < 6; ++k) {
        double x = L[k * 6 + k];
        for (int j = 0; j < k; ++j) x -= L[k * 6 + j] * L[k * 6 + j];
        if (!(x > 0.0)) return false;
        x = std::sqrt(x);
        L[k * 6 + k] = x;
        for (int i = k + 1; i < 6; ++i) {
            double s = L[i * 6 + k];
            for (int j = 0; j < k; ++j) s -= L[i * 6 + j] * L[k * 6 + j];
            L[i * 6 + k] = s / x;
        }
    }
    double inv[36];
    for (int c = 0; c < 6; ++c) {
        double v[6];
        for (int i = 0; i < 6; ++i) v[i] = (i == c) ? 1.0 : 0.0;
        for (int i = 0; i < 6; ++i) { for (int j = 0; j < i; ++j) v[i] -= L[i * 6 + j] * v[j]; v[i] /= L[i * 6 + i]; }
        for (int i = 5; i >= 0; --i) { for (int j = i + 1; j < 6; ++j) v[i] -= L[j * 6 + i] * v[j]; v[i] /= L[i * 6 + i]; }
        for (int i = 0; i < 6; ++i) inv[i * 6 + c] = v[i];
    }
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
        for (int j = 0; j < 6; ++j) s += inv[i * 6 + j] * g[j];
        y[i] = s;
    }
    return true;
}

}  // namespace

extern "C" int orc_pose_optimization(int n_obs, const double* normals, const int* levels, const double* points_w,
                                     const double pose_in[7], int max_iters, double pose_out[7], double* res_norm,
                                     orc_ba_summary* summary)
{
    const double kFunctionTol = 1e-6, kGradientTol = 1e-10, kParameterTol = 1e-8, kMinRelDecrease = 1e-3;
    const double kMinDiag = 1e-6, kMaxDiag = 1e32, kMaxRadius = 1e16, kMinRadius = 1e-32;
    const int kMaxInvalid = 5;
    orc_ba_summary sm;
    std::memset(&sm, 0, sizeof sm);

    // ref: src/Optimizer.cpp:34-36
    const Se3 T0 = se3_from(pose_in);
    double x[6] = { T0.t[0], T0.t[1], T0.t[2], 0, 0, 0 };
    so3_log(T0.q, x + 3);

    PoseBA ba;
    ba.n = n_obs; ba.normals = normals; ba.levels = levels; ba.points = points_w;
    ba.J.resize(12 * (size_t)std::max(n_obs, 1)); ba.r.resize(2 * (size_t)std::max(n_obs, 1));

    auto finish = [&](int term) {
        sm.termination = term;
        // ref: src/Optimizer.cpp:79
        Se3 T; T.q = so3_exp(x + 3); T.t[0] = x[0]; T.t[1] = x[1]; T.t[2] = x[2];
        se3_to(T, pose_out);
        // ref: src/Optimizer.cpp:298-318 GetReprojectReidual: raw residual norm of every block at the final parameters
        if (res_norm)
            for (int k = 0; k < n_obs; ++k) {
                double res[2];
                ba.block(T.q, x, k, res, nullptr);
                res_norm[k] = std::sqrt(res[0] * res[0] + res[1] * res[1]);
            }
        if (summary) *summary = sm;
        return term;
    };

    if (n_obs == 0) return finish(ORC_BA_NO_RESIDUALS);   // Ceres: "No non-constant parameter blocks found" => parameters untouched

    // ---- IterationZero
    double x_cost = 0.0;
    if (!ba.evaluate(x, true, &x_cost)) { sm.initial_cost = sm.final_cost = x_cost; return finish(ORC_BA_FAILURE); }
    sm.initial_cost = sm.final_cost = x_cost;
    double scale[6];
    for (int c = 0; c < 6; ++c) {
        double s = 0.0;
        for (int k = 0; k < 2 * n_obs; ++k) s += ba.J[6 * (size_t)k + c] * ba.J[6 * (size_t)k + c];
        scale[c] = 1.0 / (1.0 + std::sqrt(s));
    }
    auto scale_columns = [&]() { for (int k = 0; k < 2 * n_obs; ++k) for (int c = 0; c < 6; ++c) ba.J[6 * (size_t)k + c] *= scale[c]; };
    scale_columns();
    auto gradient_max_norm = [&]() {
        double ng[6], proj[6];
        for (int i = 0; i < 6; ++i) ng[i] = -ba.grad[i];
        pose_plus(x, ng, proj);
        double m = 0.0;
        for (int i = 0; i < 6; ++i) m = std::max(m, std::fabs(x[i] - proj[i]));
        return m;
    };
    double gmax = gradient_max_norm();
    double x_norm = 0.0; for (int i = 0; i < 6; ++i) x_norm += x[i] * x[i]; x_norm = std::sqrt(x_norm);

    double radius = 1e4, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    double diagonal[6];
    int n_invalid = 0;
    bool last_successful = true;   // IterationZero ends with step_is_successful = true
    int iteration = 0;

    for (;;) {
        // ---- FinalizeIterationAndCheckIfMinimizerCanContinue
        if (iteration >= max_iters) return finish(ORC_BA_NO_CONVERGENCE);
        if (last_successful && gmax <= kGradientTol) return finish(ORC_BA_GRADIENT_TOL);
        if (radius < kMinRadius) return finish(ORC_BA_MIN_RADIUS);
        ++iteration;
        sm.iterations = iteration;
        last_successful = false;

        // ---- LevenbergMarquardtStrategy::ComputeStep
        if (!reuse_diagonal) {
            for (int c = 0; c < 6; ++c) {
                double s = 0.0;
                for (int k = 0; k < 2 * n_obs; ++k) s += ba.J[6 * (size_t)k + c] * ba.J[6 * (size_t)k + c];
                diagonal[c] = std::min(std::max(s, kMinDiag), kMaxDiag);
            }
        }
        double D[6];
        for (int c = 0; c < 6; ++c) D[c] = std::sqrt(diagonal[c] / radius);
        reuse_diagonal = true;
        // SchurEliminator::BackSubstitute with the single e-block: ete = D^2 + sum e'e ; y = ete^-1 sum e'b
        double ete[36], g[6];
        for (int i = 0; i < 36; ++i) ete[i] = 0.0;
        for (int c = 0; c < 6; ++c) { ete[7 * c] = D[c] * D[c]; g[c] = 0.0; }
        for (int k = 0; k < n_obs; ++k) {
            const double* e = &ba.J[12 * (size_t)k];
            for (int c = 0; c < 6; ++c) {
                double tmp = 0.0;
                tmp += e[c] * ba.r[2 * k];
                tmp += e[6 + c] * ba.r[2 * k + 1];
                g[c] += tmp;
            }
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) {
                    double tmp = 0.0;
                    tmp += e[a] * e[b];
                    tmp += e[6 + a] * e[6 + b];
                    ete[a * 6 + b] += tmp;
                }
        }
        double step[6];
        bool valid = llt6_inverse_times(ete, g, step);
        for (int c = 0; c < 6; ++c) { if (!std::isfinite(step[c])) valid = false; step[c] = -step[c]; }
        double model_cost_change = 0.0;
        if (valid) {
            // model_cost_change = -(J step)'(f + J step / 2)
            double acc = 0.0;
            for (int k = 0; k < 2 * n_obs; ++k) {
                double m = 0.0;
                for (int c = 0; c < 6; ++c) m += ba.J[6 * (size_t)k + c] * step[c];
                acc += m * (ba.r[k] + m / 2.0);
            }
            model_cost_change = -acc;
            valid = model_cost_change > 0.0;
        }
        if (!valid) {
            // HandleInvalidStep
            if (++n_invalid >= kMaxInvalid) return finish(ORC_BA_FAILURE);
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
            continue;
        }
        n_invalid = 0;
        double delta[6], cand[6];
        for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];
        pose_plus(x, delta, cand);
        double cand_cost = 0.0;
        if (!ba.evaluate(cand, false, &cand_cost)) cand_cost = std::numeric_limits<double>::max();

        // ---- ParameterToleranceReached / FunctionToleranceReached (the candidate is NOT taken when they fire)
        double step_norm = 0.0; for (int c = 0; c < 6; ++c) step_norm += (x[c] - cand[c]) * (x[c] - cand[c]); step_norm = std::sqrt(step_norm);
        if (step_norm <= kParameterTol * (x_norm + kParameterTol)) return finish(ORC_BA_PARAMETER_TOL);
        const double cost_change = x_cost - cand_cost;
        if (std::fabs(cost_change) <= kFunctionTol * x_cost) return finish(ORC_BA_FUNCTION_TOL);

        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > kMinRelDecrease) {
            // HandleSuccessfulStep
            for (int c = 0; c < 6; ++c) x[c] = cand[c];
            x_norm = 0.0; for (int i = 0; i < 6; ++i) x_norm += x[i] * x[i]; x_norm = std::sqrt(x_norm);
            if (!ba.evaluate(x, true, &x_cost)) return finish(ORC_BA_FAILURE);
            scale_columns();
            gmax = gradient_max_norm();
            sm.final_cost = x_cost;
            sm.n_successful++;
            last_successful = true;
            const double q = 2.0 * relative_decrease - 1.0;
            radius = radius / std::max(1.0 / 3.0, 1.0 - q * q * q);
            radius = std::min(kMaxRadius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
        } else {
            radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Batched CPU driver (cpu_baseline / --impl reference arm only). One "pair" = pyramid(cur) + Run + patches.
// ------------------------------------------------------------------------------------------
int orc_pair_batch(const orc_cam* cam, int levels, const uint8_t* ref_pyrs, const uint8_t* cur_imgs, int n_pairs,
                   const orc_ref_feat* feats, int feats_per_pair, const int* n_feats, const double* ref_centers,
                   const double* poses_in, int max_level, int min_level, int max_iters, const uint8_t* patches10,
                   const double* patch_px, const int* patch_level, int patches_per_pair, int align_iters, int n_threads,
                   double* poses_out, int* n_tracked, double* patch_px_out, uint8_t* patch_conv)
{
    std::vector<int> offs(levels), ws(levels), hs(levels);
    {
        int off = 0;
        for (int l = 0; l < levels; ++l) {
            ws[l] = l ? (ws[l - 1] + 1) / 2 : cam->width;
            hs[l] = l ? (hs[l - 1] + 1) / 2 : cam->height;
            offs[l] = off; off += ws[l] * hs[l];
        }
    }
    const size_t pyr_bytes = (size_t)offs[levels - 1] + (size_t)ws[levels - 1] * hs[levels - 1];
    const size_t img_bytes = (size_t)cam->width * cam->height;
    auto work = [&](int t) {
        std::vector<uint8_t> cur(pyr_bytes);
        std::vector<int> o(levels), w(levels), h(levels);
        for (int i = t; i < n_pairs; i += n_threads) {
            orc_pyramid(cur_imgs + (size_t)i * img_bytes, cam->width, cam->height, levels, cur.data(), o.data(), w.data(), h.data());
            int nl = 0;
            n_tracked[i] = orc_sparse_align(cam, ref_pyrs + (size_t)i * pyr_bytes, cur.data(), offs.data(), ws.data(), hs.data(),
                                            feats + (size_t)i * feats_per_pair, n_feats[i], ref_centers + 3 * (size_t)i,
                                            poses_in + 7 * (size_t)i, max_level, min_level, max_iters, poses_out + 7 * (size_t)i,
                                            nullptr, 0, &nl);
            for (int p = 0; p < patches_per_pair; ++p) {
                const size_t q = (size_t)i * patches_per_pair + p;
                uint8_t p8[64];
                orc_patch_no_border(patches10 + q * 100, p8);
                double px[2] = { patch_px[2 * q], patch_px[2 * q + 1] };
                const int L = patch_level[q];
                patch_conv[q] = (uint8_t)orc_align2d(cur.data() + offs[L], ws[L], hs[L], ws[L], patches10 + q * 100, p8, align_iters, px, nullptr);
                patch_px_out[2 * q] = px[0]; patch_px_out[2 * q + 1] = px[1];
            }
        }
    };
    if (n_threads <= 1) { n_threads = 1; work(0); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    return 0;
}

// The same step with the reference's own refinement chain instead of host-provided patches (ref: src/Tracking.cpp:219-224,257-313
// TrackWithLocalMap right after Run, with the reference frame as the one key frame of the local map; src/Feature_alignment.cpp:54-69,
// 128-158; src/MapPoint.cpp:133-174): per feature with a map point -> ReprojectPoint, Get_ClosetObs (single observation), the IsInImage
// gate, SolveAffineMatrix, GetBestSearchLevel, WarpAffine, GetPatchNoBoarder, Align2DGaussNewton. Records as dsdtm_reproj.
int orc_pair_batch_map(const orc_cam* cam, int levels, int cell_size, const uint8_t* ref_pyrs, const uint8_t* cur_imgs, int n_pairs,
                       const orc_ref_feat* feats, int feats_per_pair, const int* n_feats, const double* ref_centers,
                       const double* poses_ref, const double* poses_in, int max_level, int min_level, int max_iters,
                       int points_per_pair, int max_search_level, int align_iters, int n_threads,
                       double* poses_out, int* n_tracked, orc_reproj* reproj)
{
    std::vector<int> offs(levels), ws(levels), hs(levels);
    {
        int off = 0;
        for (int l = 0; l < levels; ++l) {
            ws[l] = l ? (ws[l - 1] + 1) / 2 : cam->width;
            hs[l] = l ? (hs[l - 1] + 1) / 2 : cam->height;
            offs[l] = off; off += ws[l] * hs[l];
        }
    }
    const size_t pyr_bytes = (size_t)offs[levels - 1] + (size_t)ws[levels - 1] * hs[levels - 1];
    const size_t img_bytes = (size_t)cam->width * cam->height;
    const int grid_cols = (cam->width + cell_size - 1) / cell_size;
    auto work = [&](int t) {
        std::vector<uint8_t> cur(pyr_bytes);
        std::vector<int> o(levels), w(levels), h(levels);
        for (int i = t; i < n_pairs; i += n_threads) {
            orc_pyramid(cur_imgs + (size_t)i * img_bytes, cam->width, cam->height, levels, cur.data(), o.data(), w.data(), h.data());
            int nl = 0;
            const uint8_t* ref = ref_pyrs + (size_t)i * pyr_bytes;
            n_tracked[i] = orc_sparse_align(cam, ref, cur.data(), offs.data(), ws.data(), hs.data(),
                                            feats + (size_t)i * feats_per_pair, n_feats[i], ref_centers + 3 * (size_t)i,
                                            poses_in + 7 * (size_t)i, max_level, min_level, max_iters, poses_out + 7 * (size_t)i,
                                            nullptr, 0, &nl);
            // cur.Set_Pose(T_c2r * ref.Get_Pose()) (ref: src/Sprase_ImageAlign.cpp:57); mOw (ref: src/Frame.cpp:167-174)
            double pose_cur[7], inv[7], Tck[7];
            orc_se3_mul(poses_out + 7 * (size_t)i, poses_ref + 7 * (size_t)i, pose_cur);
            orc_se3_inv(pose_cur, inv);
            const double cur_center[3] = { inv[4], inv[5], inv[6] };
            orc_se3_inv(poses_ref + 7 * (size_t)i, inv);
            orc_se3_mul(pose_cur, inv, Tck);                                          // ref: src/Feature_alignment.cpp:181
            const double* kfc = ref_centers + 3 * (size_t)i;
            for (int j = 0; j < points_per_pair; ++j) {
                orc_reproj& r = reproj[(size_t)i * points_per_pair + j];
                r.px_proj[0] = r.px_proj[1] = r.px[0] = r.px[1] = 0.0; r.cell = -1; r.obs = -1; r.flags = 0; r.level = -1;
                if (j >= n_feats[i]) continue;
                const orc_ref_feat& f = feats[(size_t)i * feats_per_pair + j];
                if (!f.initial) continue;
                double px[2]; int cell = -1;
                if (orc_reproject_point(cam, pose_cur, f.point_w, cell_size, grid_cols, px, &cell)) { r.flags |= 1; r.cell = cell; }
                r.px_proj[0] = r.px[0] = px[0]; r.px_proj[1] = r.px[1] = px[1]; r.obs = j;
                int best = -1;
                if (orc_closest_obs(cur_center, f.point_w, kfc, 1, &best)) r.flags |= 2;
                const float sc = (float)(1 << f.level);
                if (orc_is_in_image(cam, f.px[0] / sc, f.px[1] / sc, 5, f.level)) r.flags |= 4;
                if ((r.flags & 7) != 7) continue;
                double A[4];
                orc_solve_affine(cam, kfc, f.point_w, f.normal, f.px, f.level, Tck, A);
                const int L = orc_best_search_level(A, max_search_level);
                uint8_t p10[100], p8[64];
                orc_warp_affine(A, ref + offs[f.level], ws[f.level], hs[f.level], ws[f.level], f.px, f.level, L, p10);
                orc_patch_no_border(p10, p8);
                double q[2] = { px[0] / (double)(1 << L), px[1] / (double)(1 << L) };
                if (orc_align2d(cur.data() + offs[L], ws[L], hs[L], ws[L], p10, p8, align_iters, q, nullptr)) r.flags |= 8;
                r.px[0] = q[0] * (double)(1 << L); r.px[1] = q[1] * (double)(1 << L); r.level = L;
            }
        }
    };
    if (n_threads <= 1) { n_threads = 1; work(0); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    return 0;
}

}  // extern "C"
