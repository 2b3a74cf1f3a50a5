// TEST INFRASTRUCTURE -- the few Eigen types controller.cpp / controller.hpp use, with Eigen 3.4's arithmetic
// restated: Quaternion::inverse() = conjugate / squaredNorm (zero quaternion when squaredNorm == 0),
// Quaternion * vector = v + w*(2 u x v) + u x (2 u x v) (QuaternionBase::_transformVector).  fp32, compiled with
// -ffp-contract=off like an x86-64 build of the reference without -march.
#pragma once
#include <cstddef>
#include <initializer_list>

namespace Eigen {
template <class T, int N>
struct Vector {
  T v[N];
  Vector() { for (int i = 0; i < N; ++i) v[i] = T(0); }
  Vector(std::initializer_list<T> l) { int i = 0; for (T x : l) { if (i < N) v[i++] = x; } for (; i < N; ++i) v[i] = T(0); }
  static Vector Zero() { return Vector(); }
  T& operator[](std::size_t i) { return v[i]; }
  const T& operator[](std::size_t i) const { return v[i]; }
  static Vector cross(const Vector& a, const Vector& b) {
    Vector r;
    r.v[0] = a.v[1] * b.v[2] - a.v[2] * b.v[1];
    r.v[1] = a.v[2] * b.v[0] - a.v[0] * b.v[2];
    r.v[2] = a.v[0] * b.v[1] - a.v[1] * b.v[0];
    return r;
  }
};
using Vector3f = Vector<float, 3>;

template <class V> struct Map;
template <> struct Map<const Vector3f> { const float* p; explicit Map(const float* q) : p(q) {} };
template <> struct Map<Vector3f> {
  float* p; explicit Map(float* q) : p(q) {}
  Map& operator=(const Vector3f& r) { p[0] = r.v[0]; p[1] = r.v[1]; p[2] = r.v[2]; return *this; }
};

template <class T>
struct Quaternion {
  T w_, x_, y_, z_;
  Quaternion() : w_(0), x_(0), y_(0), z_(0) {}     // Eigen leaves it uninitialised; steps are gated until the first /lowstate
  Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
  Quaternion inverse() const {
    const T n2 = ((x_ * x_ + y_ * y_) + z_ * z_) + w_ * w_;
    if (n2 > T(0)) return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
    return Quaternion(T(0), T(0), T(0), T(0));
  }
  Vector<T, 3> operator*(const Map<const Vector<T, 3>>& m) const {
    Vector<T, 3> v; v.v[0] = m.p[0]; v.v[1] = m.p[1]; v.v[2] = m.p[2];
    Vector<T, 3> u; u.v[0] = x_; u.v[1] = y_; u.v[2] = z_;
    Vector<T, 3> uv = Vector<T, 3>::cross(u, v);
    for (int i = 0; i < 3; ++i) uv.v[i] += uv.v[i];
    const Vector<T, 3> uuv = Vector<T, 3>::cross(u, uv);
    Vector<T, 3> r;
    for (int i = 0; i < 3; ++i) r.v[i] = (v.v[i] + w_ * uv.v[i]) + uuv.v[i];
    return r;
  }
};
}  // namespace Eigen
