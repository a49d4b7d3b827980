// sophus/se3.h -- TEST INFRASTRUCTURE. Stand-in for the non-templated Sophus ("1.0", ref: README.md:6; class Sophus::SE3 /
// Sophus::SO3 over Eigen::Quaterniond) that the reference includes through include/Camera.h:33. The library is not under
// /root/reference and not installed: its published algorithm (sophus/so3.cpp, sophus/se3.cpp; Eigen/src/Geometry/Quaternion.h
// for the quaternion product, conjugate, _transformVector, toRotationMatrix, rotation-matrix constructor) is RESTATED here.
// Known freedom that cannot be pinned without the library: Eigen's SSE2 quaternion product and 4-vector norm associate their sums
// differently from the scalar formulas below (last-bit effects on a quaternion that is re-normalised after every product).
// Nothing under dsdtm_b200/ includes this file.
#ifndef MINI_SOPHUS_SE3_H
#define MINI_SOPHUS_SE3_H

#include <cmath>

#include "../mini_eigen.h"

namespace Sophus {

typedef Eigen::Matrix<double, 6, 1> Vector6d;
typedef Eigen::Matrix<double, 6, 6> Matrix6d;
const double SMALL_EPS = 1e-10;

struct Quaterniond {  // Eigen::Quaterniond: coefficients stored x, y, z, w; constructor order (w, x, y, z)
    double qw, qx, qy, qz;
    Quaterniond() : qw(1), qx(0), qy(0), qz(0) {}
    Quaterniond(double w_, double x_, double y_, double z_) : qw(w_), qx(x_), qy(y_), qz(z_) {}
    explicit Quaterniond(const Eigen::Matrix3d& m)  // Eigen quaternionbase_assign_impl<Matrix3, 3, 3>
    {
        double t = m.trace();
        if (t > 0.0) {
            t = std::sqrt(t + 1.0);
            qw = 0.5 * t;
            t = 0.5 / t;
            qx = (m(2, 1) - m(1, 2)) * t;
            qy = (m(0, 2) - m(2, 0)) * t;
            qz = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            double c[3];
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
            c[i] = 0.5 * t;
            t = 0.5 / t;
            qw = (m(k, j) - m(j, k)) * t;
            c[j] = (m(j, i) + m(i, j)) * t;
            c[k] = (m(k, i) + m(i, k)) * t;
            qx = c[0]; qy = c[1]; qz = c[2];
        }
    }
    double w() const { return qw; }
    double x() const { return qx; }
    double y() const { return qy; }
    double z() const { return qz; }
    Eigen::Vector3d vec() const { return Eigen::Vector3d(qx, qy, qz); }
    double squaredNorm() const { return qx * qx + qy * qy + qz * qz + qw * qw; }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const double n = norm(); qx /= n; qy /= n; qz /= n; qw /= n; }  // coeffs() /= norm()
    void setIdentity() { qw = 1; qx = qy = qz = 0; }
    Quaterniond conjugate() const { return Quaterniond(qw, -qx, -qy, -qz); }
    Quaterniond operator*(const Quaterniond& b) const  // Eigen quat_product (generic)
    {
        const Quaterniond& a = *this;
        return Quaterniond(a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz,
                           a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy,
                           a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz,
                           a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx);
    }
    Quaterniond& operator*=(const Quaterniond& b) { *this = *this * b; return *this; }
    Eigen::Vector3d _transformVector(const Eigen::Vector3d& v) const  // uv = vec x v; uv += uv; v + w*uv + vec x uv
    {
        double uv[3] = { qy * v(2) - qz * v(1), qz * v(0) - qx * v(2), qx * v(1) - qy * v(0) };
        uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
        const double c[3] = { qy * uv[2] - qz * uv[1], qz * uv[0] - qx * uv[2], qx * uv[1] - qy * uv[0] };
        return Eigen::Vector3d(v(0) + qw * uv[0] + c[0], v(1) + qw * uv[1] + c[1], v(2) + qw * uv[2] + c[2]);
    }
    Eigen::Matrix3d toRotationMatrix() const
    {
        const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
        const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
        const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        Eigen::Matrix3d r;
        r(0, 0) = 1.0 - (tyy + tzz); r(0, 1) = txy - twz;         r(0, 2) = txz + twy;
        r(1, 0) = txy + twz;         r(1, 1) = 1.0 - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy;         r(2, 1) = tyz + twx;         r(2, 2) = 1.0 - (txx + tyy);
        return r;
    }
};

class SO3 {
public:
    SO3() { unit_quaternion_.setIdentity(); }
    SO3(const SO3& o) : unit_quaternion_(o.unit_quaternion_) {}
    explicit SO3(const Eigen::Matrix3d& R) : unit_quaternion_(R) {}
    explicit SO3(const Quaterniond& q) : unit_quaternion_(q) { unit_quaternion_.normalize(); }
    // SHIM ONLY (not Sophus API): take a unit quaternion bit for bit, so that a test can hand the reference the same 7 doubles
    // the oracle gets
    static SO3 fromUnitQuaternionRaw(const Quaterniond& q) { SO3 s; s.unit_quaternion_ = q; return s; }

    void operator=(const SO3& o) { unit_quaternion_ = o.unit_quaternion_; }
    SO3 operator*(const SO3& o) const { SO3 r(*this); r *= o; return r; }
    void operator*=(const SO3& o) { unit_quaternion_ *= o.unit_quaternion_; unit_quaternion_.normalize(); }
    Eigen::Vector3d operator*(const Eigen::Vector3d& xyz) const { return unit_quaternion_._transformVector(xyz); }
    SO3 inverse() const { return SO3(unit_quaternion_.conjugate()); }
    Eigen::Matrix3d matrix() const { return unit_quaternion_.toRotationMatrix(); }
    const Quaterniond& unit_quaternion() const { return unit_quaternion_; }

    static SO3 exp(const Eigen::Vector3d& omega) { double theta; return expAndTheta(omega, &theta); }
    static SO3 expAndTheta(const Eigen::Vector3d& omega, double* theta)
    {
        *theta = omega.norm();
        const double half_theta = 0.5 * (*theta);
        double imag_factor;
        const double real_factor = std::cos(half_theta);
        if ((*theta) < SMALL_EPS) {
            const double theta_sq = (*theta) * (*theta);
            const double theta_po4 = theta_sq * theta_sq;
            imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
        } else {
            const double sin_half_theta = std::sin(half_theta);
            imag_factor = sin_half_theta / (*theta);
        }
        return SO3(Quaterniond(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z()));
    }
    Eigen::Vector3d log() const { double theta; return logAndTheta(*this, &theta); }
    static Eigen::Vector3d logAndTheta(const SO3& other, double* theta)
    {
        const double n = other.unit_quaternion_.vec().norm();
        const double w = other.unit_quaternion_.w();
        const double squared_w = w * w;
        double two_atan_nbyw_by_n;
        if (n < SMALL_EPS) {
            two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * squared_w);
        } else {
            if (std::fabs(w) < SMALL_EPS) {
                if (w > 0) two_atan_nbyw_by_n = M_PI / n;
                else two_atan_nbyw_by_n = -M_PI / n;
            } else {
                two_atan_nbyw_by_n = 2 * std::atan(n / w) / n;
            }
        }
        *theta = two_atan_nbyw_by_n * n;
        return two_atan_nbyw_by_n * other.unit_quaternion_.vec();
    }
    static Eigen::Matrix3d hat(const Eigen::Vector3d& v)
    {
        Eigen::Matrix3d O;
        O << 0, -v(2), v(1),
             v(2), 0, -v(0),
             -v(1), v(0), 0;
        return O;
    }

private:
    Quaterniond unit_quaternion_;
};

class SE3 {
public:
    SE3() { translation_.setZero(); }
    SE3(const SO3& so3, const Eigen::Vector3d& translation) : so3_(so3), translation_(translation) {}
    SE3(const Eigen::Matrix3d& rotation_matrix, const Eigen::Vector3d& translation) : so3_(rotation_matrix), translation_(translation) {}
    SE3(const Quaterniond& quaternion, const Eigen::Vector3d& translation) : so3_(quaternion), translation_(translation) {}
    SE3(const SE3& o) : so3_(o.so3_), translation_(o.translation_) {}

    SE3& operator=(const SE3& o) { so3_ = o.so3_; translation_ = o.translation_; return *this; }
    SE3 operator*(const SE3& o) const { SE3 r(*this); r *= o; return r; }
    SE3& operator*=(const SE3& o)
    {
        translation_ += so3_ * (o.translation_);
        so3_ *= o.so3_;
        return *this;
    }
    SE3 inverse() const
    {
        SE3 ret;
        ret.so3_ = so3_.inverse();
        ret.translation_ = ret.so3_ * (translation_ * -1.);
        return ret;
    }
    Eigen::Vector3d operator*(const Eigen::Vector3d& xyz) const { return so3_ * xyz + translation_; }
    Eigen::Matrix3d rotation_matrix() const { return so3_.matrix(); }
    const Eigen::Vector3d& translation() const { return translation_; }
    Eigen::Vector3d& translation() { return translation_; }
    const SO3& so3() const { return so3_; }
    SO3& so3() { return so3_; }
    const Quaterniond& unit_quaternion() const { return so3_.unit_quaternion(); }

    static SE3 exp(const Vector6d& update)
    {
        const Eigen::Vector3d upsilon = update.head<3>();
        const Eigen::Vector3d omega = update.tail<3>();
        double theta;
        const SO3 so3 = SO3::expAndTheta(omega, &theta);
        const Eigen::Matrix3d Omega = SO3::hat(omega);
        const Eigen::Matrix3d Omega_sq = Omega * Omega;
        Eigen::Matrix3d V;
        if (theta < SMALL_EPS) {
            V = so3.matrix();
        } else {
            const double theta_sq = theta * theta;
            V = (Eigen::Matrix3d::Identity() + (1 - std::cos(theta)) / (theta_sq)*Omega + (theta - std::sin(theta)) / (theta_sq * theta) * Omega_sq);
        }
        return SE3(so3, V * upsilon);
    }

private:
    SO3 so3_;
    Eigen::Vector3d translation_;
};

}  // namespace Sophus

#endif
