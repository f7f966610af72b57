// oracle/shim/gvs_mini_eigen.hpp — TEST INFRASTRUCTURE ONLY.  Fixed-size double matrices with the few operations the
// reference's matcher uses: construction from coefficients, (i,j) / (i) access, block<R,C>(i,j) read views and
// assignment, +, -, *, norm(), Identity(), Zero(), and inverse() of square matrices (Gauss-Jordan with partial pivoting in
// double: correct to rounding, not bit-identical to Eigen's cofactor kernels -- tests therefore hand the SAME inverted
// matrices to both sides, see ref_mat4_inverse in oracle/ref_driver.cpp).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstddef>
#include <initializer_list>

namespace Eigen {

template <class T, int R, int C>
class Matrix {
public:
    T v[R * C];  // column major, as Eigen's default
    Matrix() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); v[0] = a; v[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); v[0] = a; v[1] = b; v[2] = c; }
    Matrix(T a, T b, T c, T d) { static_assert(R * C == 4, "size"); v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    template <class U>
    Matrix(std::initializer_list<U> l) {
        int i = 0;
        for (U x : l) v[i++] = (T)x;
        for (; i < R * C; ++i) v[i] = T(0);
    }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    static Matrix Zero() { return Matrix(); }
    T& operator()(int r, int c) { return v[c * R + r]; }
    const T& operator()(int r, int c) const { return v[c * R + r]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& x() { return v[0]; } T& y() { return v[1]; } T& z() { return v[2]; }
    const T& x() const { return v[0]; } const T& y() const { return v[1]; } const T& z() const { return v[2]; }

    template <int BR, int BC>
    struct Block {
        Matrix* m; int r0, c0;
        operator Matrix<T, BR, BC>() const {
            Matrix<T, BR, BC> o;
            for (int c = 0; c < BC; ++c) for (int r = 0; r < BR; ++r) o(r, c) = (*m)(r0 + r, c0 + c);
            return o;
        }
        Block& operator=(const Matrix<T, BR, BC>& o) {
            for (int c = 0; c < BC; ++c) for (int r = 0; r < BR; ++r) (*m)(r0 + r, c0 + c) = o(r, c);
            return *this;
        }
        Matrix<T, BR, BC> eval() const { return *this; }
        Matrix<T, BR, BC> operator-(const Matrix<T, BR, BC>& o) const { return eval() - o; }
        Matrix<T, BR, BC> operator+(const Matrix<T, BR, BC>& o) const { return eval() + o; }
        template <int OR_, int OC_>
        Matrix<T, BR, BC> operator-(const typename Matrix<T, OR_, OC_>::template Block<BR, BC>& o) const { return eval() - o.eval(); }
        T norm() const { return eval().norm(); }
    };
    template <int BR, int BC>
    Block<BR, BC> block(int r0, int c0) { return Block<BR, BC>{this, r0, c0}; }
    template <int BR, int BC>
    Matrix<T, BR, BC> block(int r0, int c0) const {
        Matrix<T, BR, BC> o;
        for (int c = 0; c < BC; ++c) for (int r = 0; r < BR; ++r) o(r, c) = (*this)(r0 + r, c0 + c);
        return o;
    }

    Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] + o.v[i]; return m; }
    Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] - o.v[i]; return m; }
    Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] * s; return m; }
    Matrix operator/(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.v[i] = v[i] / s; return m; }
    template <int K>
    Matrix<T, R, K> operator*(const Matrix<T, C, K>& o) const {
        Matrix<T, R, K> m;
        for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) { T s = T(0); for (int c = 0; c < C; ++c) s += (*this)(r, c) * o(c, k); m(r, k) = s; }
        return m;
    }
    Matrix inverse() const {
        static_assert(R == C, "inverse of a square matrix");
        T a[R][2 * R];
        for (int r = 0; r < R; ++r) for (int c = 0; c < R; ++c) { a[r][c] = (*this)(r, c); a[r][R + c] = r == c ? T(1) : T(0); }
        for (int p = 0; p < R; ++p) {
            int best = p;
            for (int r = p + 1; r < R; ++r) if (std::fabs(a[r][p]) > std::fabs(a[best][p])) best = r;
            if (best != p) for (int c = 0; c < 2 * R; ++c) { const T t = a[p][c]; a[p][c] = a[best][c]; a[best][c] = t; }
            const T d = a[p][p];
            for (int c = 0; c < 2 * R; ++c) a[p][c] /= d;
            for (int r = 0; r < R; ++r) {
                if (r == p) continue;
                const T f = a[r][p];
                if (f != T(0)) for (int c = 0; c < 2 * R; ++c) a[r][c] -= f * a[p][c];
            }
        }
        Matrix m;
        for (int r = 0; r < R; ++r) for (int c = 0; c < R; ++c) m(r, c) = a[r][R + c];
        return m;
    }
    T trace() const { T s = T(0); for (int i = 0; i < (R < C ? R : C); ++i) s += (*this)(i, i); return s; }
    T squaredNorm() const { T s = T(0); for (int i = 0; i < R * C; ++i) s += v[i] * v[i]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> m; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) m(c, r) = (*this)(r, c); return m; }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 3, 4> Matrix34d;
// Declared so that the reference's KeyFrame.h / Conversions.h parse; nothing on the matcher path executes them.
struct MatrixXd {};
struct Quaterniond {
    double w_, x_, y_, z_;
    Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
    explicit Quaterniond(const Matrix3d&) { std::abort(); }
    Vector4d coeffs() const { return Vector4d(x_, y_, z_, w_); }
    Quaterniond normalized() const { std::abort(); }
    Matrix3d toRotationMatrix() const { std::abort(); }
};
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;

}  // namespace Eigen
