// mini_eigen.h -- TEST INFRASTRUCTURE. A small, eager, dense-matrix stand-in for the subset of Eigen 3.2 that the
// reference's hot-path translation units use, so that those files compile UNMODIFIED from /root/reference into
// oracle/_ref/libdsdtm_ref.so (Eigen itself is not installed in this image and there is no network).
//
// What is the reference's and what is the shim's:
//   * every expression the reference writes coefficient-wise (a*x + b*y, (M - N).cwiseProduct(..), J*J^T, scalar scaling,
//     casts, floor, comparisons, indexing, storage order of .data()) is evaluated here with exactly one IEEE operation per
//     source-level operation, in source order -- these results are the reference's;
//   * what Eigen computes INSIDE a library call is restated: the association order of reductions (norm, dot, sum, products'
//     inner sums: left to right here; -DMINI_EIGEN_TREE_REDUX=1 selects the halving order of Eigen's fixed-size unroller
//     for comparison), 2x2 / 3x3 inverse (Eigen's cofactor formulas), LDLT (pivoted, lower, in place) and its solve.
// Nothing under dsdtm_b200/ includes this file.
#ifndef MINI_EIGEN_H
#define MINI_EIGEN_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <type_traits>
#include <vector>

#ifndef MINI_EIGEN_TREE_REDUX
#define MINI_EIGEN_TREE_REDUX 0
#endif

namespace Eigen {

typedef std::ptrdiff_t DenseIndex;
typedef DenseIndex Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum NoChange_t { NoChange };

namespace internal {
template <class T> struct nd { typedef T type; };
template <class D> struct traits;
inline void fail(const char* what)
{
    std::fprintf(stderr, "mini_eigen: %s\n", what);
    std::abort();
}
inline void check(bool ok, const char* what) { if (!ok) fail(what); }
template <int A, int B> struct pick { enum { value = (A != Dynamic) ? A : B }; };

// reduction of v[0..n) with `f`; fixed = the expression has a compile-time size (Eigen unrolls those)
template <class S, class F> S redux(const std::vector<S>& v, F f, bool fixed)
{
    check(!v.empty(), "reduction of an empty expression");
#if MINI_EIGEN_TREE_REDUX
    if (fixed) {
        struct R {
            static S run(const std::vector<S>& v, size_t s, size_t len, F f)
            {
                if (len == 1) return v[s];
                size_t half = len / 2;
                return f(run(v, s, half, f), run(v, s + half, len - half, f));
            }
        };
        return R::run(v, 0, v.size(), f);
    }
#else
    (void)fixed;
#endif
    S r = v[0];
    for (size_t i = 1; i < v.size(); ++i) r = f(r, v[i]);
    return r;
}
template <class S> struct add_op { S operator()(S a, S b) const { return a + b; } };
template <class S> struct max_op { S operator()(S a, S b) const { return a < b ? b : a; } };
}  // namespace internal

template <class S, int R, int C, int O = ((R == 1 && C != 1) ? RowMajor : ColMajor), int MR = R, int MC = C> class Matrix;
template <class M, int BR, int BC> class Block;
template <class M> class Map;
template <class M> class LDLT;
template <class D> class ColwiseOp;
template <class M> class ArrayWrap;

// ------------------------------------------------------------------------------------------------------------------------
// Base: everything that only reads. Every operation evaluates eagerly into a plain Matrix.
// ------------------------------------------------------------------------------------------------------------------------
template <class D> class Base {
public:
    typedef typename internal::traits<D>::Scalar Scalar;
    enum { RowsAtCompileTime = internal::traits<D>::Rows, ColsAtCompileTime = internal::traits<D>::Cols,
           IsFixed = (RowsAtCompileTime != Dynamic && ColsAtCompileTime != Dynamic) };
    typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> Plain;

    const D& derived() const { return *static_cast<const D*>(this); }
    D& derived() { return *static_cast<D*>(this); }
    Index rows() const { return derived().rows(); }
    Index cols() const { return derived().cols(); }
    Index size() const { return rows() * cols(); }
    Scalar coeff(Index i, Index j) const { return derived().coeff(i, j); }
    Scalar coeff(Index k) const
    {
        if (rows() == 1) return coeff(0, k);
        if (cols() == 1) return coeff(k, 0);
        return derived().coeffLinear(k);
    }
    Scalar coeffLinear(Index k) const { return coeff(k % rows(), k / rows()); }
    Scalar operator()(Index i, Index j) const { return coeff(i, j); }
    Scalar operator()(Index k) const { return coeff(k); }
    Scalar operator[](Index k) const { return coeff(k); }
    Scalar x() const { return coeff(0); }
    Scalar y() const { return coeff(1); }
    Scalar z() const { return coeff(2); }

    Plain eval() const
    {
        Plain r;
        r.resizeLike(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j);
        return r;
    }

    // ---- reductions ----
    std::vector<Scalar> flat() const  // column-major walk (vectors: their natural order)
    {
        std::vector<Scalar> v;
        v.reserve((size_t)size());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) v.push_back(coeff(i, j));
        return v;
    }
    Scalar sum() const { return internal::redux(flat(), internal::add_op<Scalar>(), IsFixed); }
    Scalar squaredNorm() const
    {
        std::vector<Scalar> v = flat();
        for (size_t i = 0; i < v.size(); ++i) v[i] = v[i] * v[i];
        return internal::redux(v, internal::add_op<Scalar>(), IsFixed);
    }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    template <class E> Scalar dot(const Base<E>& o) const
    {
        internal::check(size() == o.size(), "dot: size mismatch");
        std::vector<Scalar> v;
        for (Index k = 0; k < size(); ++k) v.push_back(coeff(k) * o.coeff(k));
        return internal::redux(v, internal::add_op<Scalar>(), IsFixed);
    }
    Scalar maxCoeff() const { return internal::redux(flat(), internal::max_op<Scalar>(), IsFixed); }
    Scalar trace() const
    {
        std::vector<Scalar> v;
        for (Index k = 0; k < rows(); ++k) v.push_back(coeff(k, k));
        return internal::redux(v, internal::add_op<Scalar>(), IsFixed);
    }
    bool isZero(Scalar prec = Scalar(0)) const  // Eigen: every |coeff| <= prec
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i)
                if (std::abs(coeff(i, j)) > prec) return false;
        return true;
    }
    Scalar determinant() const
    {
        internal::check(rows() == cols() && (rows() == 2 || rows() == 3), "determinant: only 2x2 / 3x3");
        if (rows() == 2) return coeff(0, 0) * coeff(1, 1) - coeff(1, 0) * coeff(0, 1);  // Eigen determinant_impl<.,2>
        // Eigen determinant_impl<.,3>: bruteforce_det3_helper
        return coeff(0, 0) * (coeff(1, 1) * coeff(2, 2) - coeff(1, 2) * coeff(2, 1))
             + coeff(0, 1) * (coeff(1, 2) * coeff(2, 0) - coeff(1, 0) * coeff(2, 2))
             + coeff(0, 2) * (coeff(1, 0) * coeff(2, 1) - coeff(1, 1) * coeff(2, 0));
    }

    // ---- coefficient-wise results ----
    Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> transpose() const
    {
        Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> r;
        r.resizeLike(cols(), rows());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(j, i) = coeff(i, j);
        return r;
    }
    template <class T> Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const
    {
        Matrix<T, RowsAtCompileTime, ColsAtCompileTime> r;
        r.resizeLike(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = static_cast<T>(coeff(i, j));
        return r;
    }
    template <class F> Matrix<typename F::result_type, RowsAtCompileTime, ColsAtCompileTime> unaryExpr(F f) const
    {
        Matrix<typename F::result_type, RowsAtCompileTime, ColsAtCompileTime> r;
        r.resizeLike(rows(), cols());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = f(coeff(i, j));
        return r;
    }
    Plain cwiseAbs() const
    {
        Plain r = eval();
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::abs(coeff(i, j));
        return r;
    }
    template <class E> Plain cwiseProduct(const Base<E>& o) const
    {
        internal::check(rows() == o.rows() && cols() == o.cols(), "cwiseProduct: shape mismatch");
        Plain r = eval();
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j) * o.coeff(i, j);
        return r;
    }
    Plain normalized() const
    {
        Plain r = eval();
        const Scalar n = norm();
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j) / n;
        return r;
    }
    Plain inverse() const;  // 2x2 and 3x3, Eigen's cofactor formulas
    LDLT<Plain> ldlt() const;

    // ---- read-only sub-blocks: copies ----
    Matrix<Scalar, Dynamic, Dynamic> block(Index r0, Index c0, Index nr, Index nc) const
    {
        internal::check(r0 >= 0 && c0 >= 0 && r0 + nr <= rows() && c0 + nc <= cols(), "block: out of range");
        Matrix<Scalar, Dynamic, Dynamic> r(nr, nc);
        for (Index j = 0; j < nc; ++j)
            for (Index i = 0; i < nr; ++i) r.coeffRef(i, j) = coeff(r0 + i, c0 + j);
        return r;
    }
    Matrix<Scalar, RowsAtCompileTime, 1> col(Index j) const
    {
        Matrix<Scalar, RowsAtCompileTime, 1> r;
        r.resizeLike(rows(), 1);
        for (Index i = 0; i < rows(); ++i) r.coeffRef(i, 0) = coeff(i, j);
        return r;
    }
    Matrix<Scalar, 1, ColsAtCompileTime> row(Index i) const
    {
        Matrix<Scalar, 1, ColsAtCompileTime> r;
        r.resizeLike(1, cols());
        for (Index j = 0; j < cols(); ++j) r.coeffRef(0, j) = coeff(i, j);
        return r;
    }
    template <int N> Matrix<Scalar, N, 1> head() const
    {
        internal::check(N <= size(), "head: out of range");
        Matrix<Scalar, N, 1> r;
        for (Index k = 0; k < N; ++k) r.coeffRef(k, 0) = coeff(k);
        return r;
    }
    template <int N> Matrix<Scalar, N, 1> tail() const
    {
        internal::check(N <= size(), "tail: out of range");
        Matrix<Scalar, N, 1> r;
        for (Index k = 0; k < N; ++k) r.coeffRef(k, 0) = coeff(size() - N + k);
        return r;
    }
    ArrayWrap<Plain> array() const;
    ColwiseOp<D> colwise() const;
    ColwiseOp<D> rowwise() const;
};

// ---- arithmetic on expressions (free functions; the scalar operand is in a non-deduced context, like Eigen's
//      operator*(const Scalar&): an int or float argument converts to the matrix's scalar type) ----
#define MINI_EIGEN_BIN(op)                                                                                               \
    template <class A, class B>                                                                                          \
    Matrix<typename internal::traits<A>::Scalar, internal::pick<internal::traits<A>::Rows, internal::traits<B>::Rows>::value, \
           internal::pick<internal::traits<A>::Cols, internal::traits<B>::Cols>::value>                                  \
    operator op(const Base<A>& a, const Base<B>& b)                                                                      \
    {                                                                                                                    \
        static_assert(std::is_same<typename internal::traits<A>::Scalar, typename internal::traits<B>::Scalar>::value,  \
                      "mini_eigen: mixed scalar types");                                                                 \
        internal::check(a.rows() == b.rows() && a.cols() == b.cols(), "binary op: shape mismatch");                      \
        Matrix<typename internal::traits<A>::Scalar, internal::pick<internal::traits<A>::Rows, internal::traits<B>::Rows>::value, \
               internal::pick<internal::traits<A>::Cols, internal::traits<B>::Cols>::value> r;                           \
        r.resizeLike(a.rows(), a.cols());                                                                                \
        for (Index j = 0; j < a.cols(); ++j)                                                                             \
            for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) op b.coeff(i, j);                      \
        return r;                                                                                                        \
    }
MINI_EIGEN_BIN(+)
MINI_EIGEN_BIN(-)
#undef MINI_EIGEN_BIN

template <class A> typename Base<A>::Plain operator-(const Base<A>& a)
{
    typename Base<A>::Plain r = a.eval();
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = -a.coeff(i, j);
    return r;
}
template <class A> typename Base<A>::Plain operator*(const Base<A>& a, typename internal::nd<typename internal::traits<A>::Scalar>::type s)
{
    typename Base<A>::Plain r = a.eval();
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) * s;
    return r;
}
template <class A> typename Base<A>::Plain operator*(typename internal::nd<typename internal::traits<A>::Scalar>::type s, const Base<A>& a)
{
    typename Base<A>::Plain r = a.eval();
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = s * a.coeff(i, j);
    return r;
}
// Eigen 3.2 scalar_quotient1_op: a / s (a true division per coefficient)
template <class A> typename Base<A>::Plain operator/(const Base<A>& a, typename internal::nd<typename internal::traits<A>::Scalar>::type s)
{
    typename Base<A>::Plain r = a.eval();
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) / s;
    return r;
}
// matrix product: coefficient (i,j) = sum over k, left to right (Eigen's coefficient-based product for small fixed sizes,
// and the order in which its GEMM accumulates one coefficient)
template <class A, class B>
Matrix<typename internal::traits<A>::Scalar, internal::traits<A>::Rows, internal::traits<B>::Cols> operator*(const Base<A>& a, const Base<B>& b)
{
    static_assert(std::is_same<typename internal::traits<A>::Scalar, typename internal::traits<B>::Scalar>::value, "mini_eigen: mixed scalar types");
    typedef typename internal::traits<A>::Scalar S;
    internal::check(a.cols() == b.rows(), "product: inner dimension mismatch");
    Matrix<S, internal::traits<A>::Rows, internal::traits<B>::Cols> r;
    r.resizeLike(a.rows(), b.cols());
    for (Index j = 0; j < b.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) {
            internal::check(a.cols() > 0, "product: empty inner dimension");
            S s = a.coeff(i, 0) * b.coeff(0, j);
            for (Index k = 1; k < a.cols(); ++k) s += a.coeff(i, k) * b.coeff(k, j);
            r.coeffRef(i, j) = s;
        }
    return r;
}

template <class D> std::ostream& operator<<(std::ostream& os, const Base<D>& m)
{
    for (Index i = 0; i < m.rows(); ++i) {
        for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m.coeff(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

// ------------------------------------------------------------------------------------------------------------------------
// comma initialiser: fills row by row, left to right (independent of the storage order), like Eigen
// ------------------------------------------------------------------------------------------------------------------------
template <class T> class CommaInit {
public:
    CommaInit(T& t, typename T::Scalar first) : t_(t), k_(0) { put(first); }
    CommaInit& operator,(typename T::Scalar v) { put(v); return *this; }
    ~CommaInit() { if (k_ != t_.rows() * t_.cols()) internal::fail("comma initialiser: wrong number of coefficients"); }
private:
    void put(typename T::Scalar v)
    {
        internal::check(k_ < t_.rows() * t_.cols(), "comma initialiser: too many coefficients");
        t_.coeffRef(k_ / t_.cols(), k_ % t_.cols()) = v;
        ++k_;
    }
    T& t_;
    Index k_;
};

// ------------------------------------------------------------------------------------------------------------------------
// WBase: what writes. D provides rows(), cols(), coeff(i,j), coeffRef(i,j) (and resizeLike for plain matrices).
// ------------------------------------------------------------------------------------------------------------------------
template <class D> class WBase : public Base<D> {
public:
    typedef typename Base<D>::Scalar Scalar;
    using Base<D>::derived;
    using Base<D>::rows;
    using Base<D>::cols;
    using Base<D>::size;
    using Base<D>::coeff;
    using Base<D>::operator();
    using Base<D>::operator[];
    using Base<D>::block;
    using Base<D>::col;
    using Base<D>::row;
    using Base<D>::x;
    using Base<D>::y;
    using Base<D>::z;

    Scalar& coeffRef(Index i, Index j) { return derived().coeffRef(i, j); }
    Scalar& coeffRef(Index k)
    {
        if (rows() == 1) return coeffRef(0, k);
        if (cols() == 1) return coeffRef(k, 0);
        return coeffRef(k % rows(), k / rows());
    }
    Scalar& operator()(Index i, Index j) { return coeffRef(i, j); }
    Scalar& operator()(Index k) { return coeffRef(k); }
    Scalar& operator[](Index k) { return coeffRef(k); }
    Scalar& x() { return coeffRef(0); }
    Scalar& y() { return coeffRef(1); }
    Scalar& z() { return coeffRef(2); }

    template <class E> void assignFrom(const Base<E>& e)
    {
        static_assert(std::is_same<Scalar, typename internal::traits<E>::Scalar>::value, "mini_eigen: assignment between scalar types needs cast<>()");
        typename Base<E>::Plain t = e.eval();  // aliasing-safe
        derived().resizeFor(t.rows(), t.cols());
        if (rows() == t.rows() && cols() == t.cols()) {
            for (Index j = 0; j < cols(); ++j)
                for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = t.coeff(i, j);
        } else if ((rows() == 1 || cols() == 1) && (t.rows() == 1 || t.cols() == 1) && size() == t.size()) {
            for (Index k = 0; k < size(); ++k) coeffRef(k) = t.coeff(k);  // vector <- transposed vector
        } else {
            internal::fail("assignment: shape mismatch");
        }
    }
    template <class E> D& operator+=(const Base<E>& e)
    {
        internal::check(rows() == e.rows() && cols() == e.cols(), "+=: shape mismatch");
        typename Base<E>::Plain t = e.eval();
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += t.coeff(i, j);
        return derived();
    }
    template <class E> D& operator-=(const Base<E>& e)
    {
        internal::check(rows() == e.rows() && cols() == e.cols(), "-=: shape mismatch");
        typename Base<E>::Plain t = e.eval();
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= t.coeff(i, j);
        return derived();
    }
    D& operator*=(Scalar s)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) *= s;
        return derived();
    }
    D& operator/=(Scalar s)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) /= s;
        return derived();
    }
    D& setConstant(Scalar v)
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = v;
        return derived();
    }
    D& setZero() { return setConstant(Scalar(0)); }
    D& setOnes() { return setConstant(Scalar(1)); }
    D& setIdentity()
    {
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0);
        return derived();
    }
    void normalize() { *this /= this->norm(); }  // Eigen 3.2: *this /= norm()
    D& noalias() { return derived(); }
    CommaInit<D> operator<<(Scalar first) { return CommaInit<D>(derived(), first); }

    // writable views
    Block<D, Dynamic, Dynamic> block(Index r0, Index c0, Index nr, Index nc) { return Block<D, Dynamic, Dynamic>(derived(), r0, c0, nr, nc); }
    Block<D, internal::traits<D>::Rows, 1> col(Index j) { return Block<D, internal::traits<D>::Rows, 1>(derived(), 0, j, rows(), 1); }
    Block<D, 1, internal::traits<D>::Cols> row(Index i) { return Block<D, 1, internal::traits<D>::Cols>(derived(), i, 0, 1, cols()); }
    template <int N> Block<D, N, 1> head()
    {
        internal::check(cols() == 1 && N <= rows(), "head<N>() on a non-column");
        return Block<D, N, 1>(derived(), 0, 0, N, 1);
    }
    template <int N> Block<D, N, 1> tail()
    {
        internal::check(cols() == 1 && N <= rows(), "tail<N>() on a non-column");
        return Block<D, N, 1>(derived(), rows() - N, 0, N, 1);
    }
    template <int N> Matrix<Scalar, N, 1> head() const { return Base<D>::template head<N>(); }
    template <int N> Matrix<Scalar, N, 1> tail() const { return Base<D>::template tail<N>(); }
};

// ------------------------------------------------------------------------------------------------------------------------
// Matrix
// ------------------------------------------------------------------------------------------------------------------------
namespace internal {
template <class S, int R, int C, int O, int MR, int MC> struct traits<Matrix<S, R, C, O, MR, MC> > {
    typedef S Scalar;
    enum { Rows = R, Cols = C };
};
}

template <class S, int R, int C, int O, int MR, int MC> class Matrix : public WBase<Matrix<S, R, C, O, MR, MC> > {
    typedef WBase<Matrix<S, R, C, O, MR, MC> > W;
public:
    typedef S Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C, Options = O, IsRowMajor = (O & RowMajor) ? 1 : 0 };
    using W::operator();
    using W::operator[];
    using W::coeff;
    using W::coeffRef;

    Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)(r_ * c_), S(0)) {}
    Matrix(const Matrix& o) : r_(o.r_), c_(o.c_), d_(o.d_) {}
    explicit Matrix(const S* p) : r_(R), c_(C), d_(p, p + (size_t)R * C) { static_assert(R != Dynamic && C != Dynamic, "pointer ctor needs a fixed size"); }
    // (rows, cols) for dynamic sizes; (x, y) for fixed 2-vectors -- like Eigen's two-argument constructor
    template <class A, class B> Matrix(const A& a, const B& b) : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C)
    {
        init2(a, b, std::integral_constant<bool, (R != Dynamic && C != Dynamic)>());
    }
    Matrix(const S& a, const S& b, const S& c) : r_(R), c_(C), d_((size_t)3)
    {
        static_assert(R != Dynamic && C != Dynamic && R * C == 3, "three-coefficient ctor needs a fixed 3-vector");
        d_[0] = a; d_[1] = b; d_[2] = c;
    }
    template <class E> Matrix(const Base<E>& e) : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)(r_ * c_), S(0)) { this->assignFrom(e); }

    Matrix& operator=(const Matrix& o)
    {
        if (this != &o) { resizeFor(o.r_, o.c_); this->assignFrom(static_cast<const Base<Matrix>&>(o)); }
        return *this;
    }
    template <class E> Matrix& operator=(const Base<E>& e) { this->assignFrom(e); return *this; }

    Index rows() const { return r_; }
    Index cols() const { return c_; }
    S coeff(Index i, Index j) const { return d_[idx(i, j)]; }
    S& coeffRef(Index i, Index j) { return d_[idx(i, j)]; }
    S coeffLinear(Index k) const { return d_[(size_t)k]; }  // storage order
    S* data() { return d_.data(); }
    const S* data() const { return d_.data(); }

    // destructive resize (contents unspecified in Eigen; zero here)
    void resize(Index r, Index c)
    {
        internal::check((R == Dynamic || r == R) && (C == Dynamic || c == C), "resize: fixed dimension changed");
        r_ = r; c_ = c;
        d_.assign((size_t)(r * c), S(0));
    }
    void resize(NoChange_t, Index c) { resize(r_, c); }
    void resize(Index r, NoChange_t) { resize(r, c_); }
    void conservativeResize(Index r, Index c)
    {
        Matrix old(*this);
        resize(r, c);
        for (Index j = 0; j < std::min(c, old.c_); ++j)
            for (Index i = 0; i < std::min(r, old.r_); ++i) coeffRef(i, j) = old.coeff(i, j);
    }
    void conservativeResize(NoChange_t, Index c) { conservativeResize(r_, c); }
    void conservativeResize(Index r, NoChange_t) { conservativeResize(r, c_); }
    // used by the shim itself: give an expression result its shape
    void resizeLike(Index r, Index c)
    {
        if (r != r_ || c != c_) resize(r, c);
    }
    void resizeFor(Index r, Index c)  // on assignment: dynamic dimensions follow the right-hand side (Eigen resizes on =)
    {
        if (r == r_ && c == c_) return;
        if ((R == Dynamic || R == r) && (C == Dynamic || C == c)) { resize(r, c); return; }
        if ((R == 1 || C == 1) && (r == 1 || c == 1)) {  // vector <- transposed vector
            const Index n = r * c;
            if (R == 1 && (C == Dynamic || C == n)) { resize(1, n); return; }
            if (C == 1 && (R == Dynamic || R == n)) { resize(n, 1); return; }
        }
    }

    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(Index r, Index c) { Matrix m; m.resize(r, c); return m; }
    static Matrix Ones() { Matrix m; m.setOnes(); return m; }
    static Matrix Ones(Index r, Index c) { Matrix m; m.resize(r, c); m.setOnes(); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(Index r, Index c) { Matrix m; m.resize(r, c); m.setIdentity(); return m; }

private:
    size_t idx(Index i, Index j) const
    {
#ifndef NDEBUG
        if (i < 0 || j < 0 || i >= r_ || j >= c_) internal::fail("coefficient index out of range");
#endif
        return IsRowMajor ? (size_t)(i * c_ + j) : (size_t)(j * r_ + i);
    }
    template <class A, class B> void init2(const A& a, const B& b, std::true_type)
    {
        static_assert(R * C == 2 || R == Dynamic, "two-coefficient ctor needs a fixed 2-vector");
        d_.assign(2, S(0));
        d_[0] = S(a); d_[1] = S(b);
    }
    template <class A, class B> void init2(const A& a, const B& b, std::false_type) { resize((Index)a, (Index)b); }

    Index r_, c_;
    std::vector<S> d_;
};

// ------------------------------------------------------------------------------------------------------------------------
// Block (writable view of a Matrix / Map / Block) and Map (view of raw memory in M's storage order)
// ------------------------------------------------------------------------------------------------------------------------
namespace internal {
template <class M, int BR, int BC> struct traits<Block<M, BR, BC> > {
    typedef typename traits<M>::Scalar Scalar;
    enum { Rows = BR, Cols = BC };
};
template <class M> struct traits<Map<M> > {
    typedef typename traits<typename std::remove_const<M>::type>::Scalar Scalar;
    enum { Rows = traits<typename std::remove_const<M>::type>::Rows, Cols = traits<typename std::remove_const<M>::type>::Cols };
};
}

template <class M, int BR, int BC> class Block : public WBase<Block<M, BR, BC> > {
    typedef WBase<Block<M, BR, BC> > W;
public:
    typedef typename internal::traits<M>::Scalar Scalar;
    using W::operator();
    using W::operator[];
    using W::coeff;
    using W::coeffRef;
    Block(M& m, Index r0, Index c0, Index nr, Index nc) : m_(&m), r0_(r0), c0_(c0), nr_(nr), nc_(nc)
    {
        internal::check(r0 >= 0 && c0 >= 0 && nr >= 0 && nc >= 0 && r0 + nr <= m.rows() && c0 + nc <= m.cols(), "block: out of range");
    }
    Block(const Block& o) : m_(o.m_), r0_(o.r0_), c0_(o.c0_), nr_(o.nr_), nc_(o.nc_) {}
    Block& operator=(const Block& o) { this->assignFrom(static_cast<const Base<Block>&>(o)); return *this; }
    template <class E> Block& operator=(const Base<E>& e) { this->assignFrom(e); return *this; }
    Index rows() const { return nr_; }
    Index cols() const { return nc_; }
    Scalar coeff(Index i, Index j) const { return static_cast<const M*>(m_)->coeff(r0_ + i, c0_ + j); }
    Scalar& coeffRef(Index i, Index j) { return m_->coeffRef(r0_ + i, c0_ + j); }
    void resizeFor(Index, Index) {}
private:
    M* m_;
    Index r0_, c0_, nr_, nc_;
};

template <class M> class Map : public WBase<Map<M> > {
    typedef WBase<Map<M> > W;
    typedef typename std::remove_const<M>::type PM;
public:
    typedef typename internal::traits<PM>::Scalar Scalar;
    typedef typename std::conditional<std::is_const<M>::value, const Scalar, Scalar>::type* Ptr;
    using W::operator();
    using W::operator[];
    using W::coeff;
    using W::coeffRef;
    explicit Map(Ptr p) : p_(const_cast<Scalar*>(p)), r_(PM::RowsAtCompileTime), c_(PM::ColsAtCompileTime)
    {
        static_assert(PM::RowsAtCompileTime != Dynamic && PM::ColsAtCompileTime != Dynamic, "Map(ptr) needs a fixed size");
    }
    Map(Ptr p, Index r, Index c) : p_(const_cast<Scalar*>(p)), r_(r), c_(c) {}
    Map(const Map& o) : p_(o.p_), r_(o.r_), c_(o.c_) {}
    Map& operator=(const Map& o) { this->assignFrom(static_cast<const Base<Map>&>(o)); return *this; }
    template <class E> Map& operator=(const Base<E>& e) { this->assignFrom(e); return *this; }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Scalar coeff(Index i, Index j) const { return p_[idx(i, j)]; }
    Scalar& coeffRef(Index i, Index j) { return p_[idx(i, j)]; }
    Scalar coeffLinear(Index k) const { return p_[k]; }
    Scalar* data() { return p_; }
    void resizeFor(Index, Index) {}
private:
    size_t idx(Index i, Index j) const { return PM::IsRowMajor ? (size_t)(i * c_ + j) : (size_t)(j * r_ + i); }
    Scalar* p_;
    Index r_, c_;
};

// ------------------------------------------------------------------------------------------------------------------------
// colwise() / rowwise() and array(): only the forms the reference uses
//   M.colwise() - v, M.colwise() + v   (v a column)        ref: src/Sprase_ImageAlign.cpp:113, src/Feature_alignment.cpp:232
//   M.colwise().norm()                                      ref: src/Sprase_ImageAlign.cpp:114
//   M.array().rowwise() * r.array()    (r a row)            ref: src/Sprase_ImageAlign.cpp:115
// ------------------------------------------------------------------------------------------------------------------------
template <class D> class ColwiseOp {
public:
    typedef typename internal::traits<D>::Scalar Scalar;
    typedef Matrix<Scalar, internal::traits<D>::Rows, internal::traits<D>::Cols> Plain;
    ColwiseOp(const D& d, bool colwise) : d_(d), colwise_(colwise) {}
    template <class E> Plain operator-(const Base<E>& v) const { return apply(v, false); }
    template <class E> Plain operator+(const Base<E>& v) const { return apply(v, true); }
    Matrix<Scalar, 1, internal::traits<D>::Cols> norm() const
    {
        internal::check(colwise_, "rowwise().norm() is not in the shim");
        Matrix<Scalar, 1, internal::traits<D>::Cols> r;
        r.resizeLike(1, d_.cols());
        for (Index j = 0; j < d_.cols(); ++j) r.coeffRef(0, j) = static_cast<const Base<D>&>(d_).col(j).norm();
        return r;
    }
    template <class M2> Plain operator*(const ArrayWrap<M2>& a) const;  // rowwise() * row-array
private:
    template <class E> Plain apply(const Base<E>& v, bool add) const
    {
        internal::check(colwise_ && v.cols() == 1 && v.rows() == d_.rows(), "colwise() +/-: needs a matching column");
        Plain r;
        r.resizeLike(d_.rows(), d_.cols());
        for (Index j = 0; j < d_.cols(); ++j)
            for (Index i = 0; i < d_.rows(); ++i)
                r.coeffRef(i, j) = add ? d_.coeff(i, j) + v.coeff(i, 0) : d_.coeff(i, j) - v.coeff(i, 0);
        return r;
    }
    const D& d_;
    bool colwise_;
};

template <class M> class ArrayWrap {
public:
    explicit ArrayWrap(const M& m) : m_(m) {}
    const M& matrix() const { return m_; }
    ColwiseOp<M> rowwise() const { return ColwiseOp<M>(m_, false); }
    ColwiseOp<M> colwise() const { return ColwiseOp<M>(m_, true); }
private:
    M m_;
};

template <class D> template <class M2> typename ColwiseOp<D>::Plain ColwiseOp<D>::operator*(const ArrayWrap<M2>& a) const
{
    const M2& v = a.matrix();
    internal::check(!colwise_ && v.rows() == 1 && v.cols() == d_.cols(), "rowwise() *: needs a matching row");
    Plain r;
    r.resizeLike(d_.rows(), d_.cols());
    for (Index j = 0; j < d_.cols(); ++j)
        for (Index i = 0; i < d_.rows(); ++i) r.coeffRef(i, j) = d_.coeff(i, j) * v.coeff(0, j);
    return r;
}

template <class D> ArrayWrap<typename Base<D>::Plain> Base<D>::array() const { return ArrayWrap<Plain>(eval()); }
template <class D> ColwiseOp<D> Base<D>::colwise() const { return ColwiseOp<D>(derived(), true); }
template <class D> ColwiseOp<D> Base<D>::rowwise() const { return ColwiseOp<D>(derived(), false); }

// ------------------------------------------------------------------------------------------------------------------------
// inverse(): Eigen/src/LU/Inverse.h, compute_inverse<.,.,2> and <.,.,3> (cofactors; not under /root/reference, restated)
// ------------------------------------------------------------------------------------------------------------------------
template <class D> typename Base<D>::Plain Base<D>::inverse() const
{
    internal::check(rows() == cols() && (rows() == 2 || rows() == 3), "inverse: only 2x2 / 3x3");
    Plain r;
    r.resizeLike(rows(), cols());
    if (rows() == 2) {
        const Scalar invdet = Scalar(1) / determinant();
        r.coeffRef(0, 0) = coeff(1, 1) * invdet;
        r.coeffRef(1, 0) = -coeff(1, 0) * invdet;
        r.coeffRef(0, 1) = -coeff(0, 1) * invdet;
        r.coeffRef(1, 1) = coeff(0, 0) * invdet;
        return r;
    }
    struct Cof {
        static Scalar at(const Base<D>& m, int i, int j)
        {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m.coeff(i1, j1) * m.coeff(i2, j2) - m.coeff(i1, j2) * m.coeff(i2, j1);
        }
    };
    Matrix<Scalar, 3, 1> c0;
    c0.coeffRef(0, 0) = Cof::at(*this, 0, 0);
    c0.coeffRef(1, 0) = Cof::at(*this, 1, 0);
    c0.coeffRef(2, 0) = Cof::at(*this, 2, 0);
    const Scalar det = c0.cwiseProduct(col(0)).sum();
    const Scalar invdet = Scalar(1) / det;
    r.coeffRef(0, 0) = c0.coeff(0, 0) * invdet;
    r.coeffRef(0, 1) = c0.coeff(1, 0) * invdet;
    r.coeffRef(0, 2) = c0.coeff(2, 0) * invdet;
    r.coeffRef(1, 0) = Cof::at(*this, 0, 1) * invdet;
    r.coeffRef(1, 1) = Cof::at(*this, 1, 1) * invdet;
    r.coeffRef(1, 2) = Cof::at(*this, 2, 1) * invdet;
    r.coeffRef(2, 0) = Cof::at(*this, 0, 2) * invdet;
    r.coeffRef(2, 1) = Cof::at(*this, 1, 2) * invdet;
    r.coeffRef(2, 2) = Cof::at(*this, 2, 2) * invdet;
    return r;
}

// ------------------------------------------------------------------------------------------------------------------------
// LDLT: Eigen/src/Cholesky/LDLT.h (3.2), ldlt_inplace<Lower>::unblocked + solve. Not under /root/reference: restated.
// Pivoting on the largest remaining diagonal entry, L D L^T on the lower triangle in place; solve = P^T L^-T D^+ L^-1 P b,
// where D^+ zeroes the components whose pivot is not above 1/highest.
// ------------------------------------------------------------------------------------------------------------------------
template <class M> class LDLT {
public:
    typedef typename M::Scalar Scalar;
    explicit LDLT(const M& a) : m_(a), n_(a.rows()), tr_((size_t)a.rows())
    {
        internal::check(a.rows() == a.cols(), "ldlt: not square");
        const Index n = n_;
        std::vector<Scalar> temp((size_t)n);
        for (Index k = 0; k < n; ++k) {
            Index big = k;
            Scalar bigv = std::abs(m_.coeff(k, k));
            for (Index i = k + 1; i < n; ++i) {
                const Scalar v = std::abs(m_.coeff(i, i));
                if (v > bigv) { bigv = v; big = i; }
            }
            tr_[(size_t)k] = big;
            if (k != big) {
                for (Index j = 0; j < k; ++j) std::swap(m_.coeffRef(k, j), m_.coeffRef(big, j));
                for (Index i = big + 1; i < n; ++i) std::swap(m_.coeffRef(i, k), m_.coeffRef(i, big));
                std::swap(m_.coeffRef(k, k), m_.coeffRef(big, big));
                for (Index i = k + 1; i < big; ++i) std::swap(m_.coeffRef(i, k), m_.coeffRef(big, i));
            }
            const Index rs = n - k - 1;
            if (k > 0) {
                for (Index j = 0; j < k; ++j) temp[(size_t)j] = m_.coeff(j, j) * m_.coeff(k, j);
                Scalar s = 0;
                for (Index j = 0; j < k; ++j) s += m_.coeff(k, j) * temp[(size_t)j];
                m_.coeffRef(k, k) -= s;
                for (Index i = k + 1; i < n; ++i) {
                    Scalar s2 = 0;
                    for (Index j = 0; j < k; ++j) s2 += m_.coeff(i, j) * temp[(size_t)j];
                    m_.coeffRef(i, k) -= s2;
                }
            }
            const Scalar akk = m_.coeff(k, k);
            const bool valid = std::abs(akk) > Scalar(0);
            if (k == 0 && !valid) {
                for (Index j = 0; j < n; ++j) tr_[(size_t)j] = j;
                break;
            }
            if (rs > 0 && valid)
                for (Index i = k + 1; i < n; ++i) m_.coeffRef(i, k) /= akk;
        }
    }
    template <class E> Matrix<Scalar, M::RowsAtCompileTime, 1> solve(const Base<E>& b) const
    {
        internal::check(b.cols() == 1 && b.rows() == n_, "ldlt.solve: needs a matching column");
        const Index n = n_;
        std::vector<Scalar> y((size_t)n);
        for (Index i = 0; i < n; ++i) y[(size_t)i] = b.coeff(i, 0);
        for (Index k = 0; k < n; ++k) if (tr_[(size_t)k] != k) std::swap(y[(size_t)k], y[(size_t)tr_[(size_t)k]]);
        for (Index i = 0; i < n; ++i)
            for (Index j = 0; j < i; ++j) y[(size_t)i] -= m_.coeff(i, j) * y[(size_t)j];
        const Scalar tol = Scalar(1) / std::numeric_limits<Scalar>::max();
        for (Index i = 0; i < n; ++i) {
            if (std::abs(m_.coeff(i, i)) > tol) y[(size_t)i] /= m_.coeff(i, i);
            else y[(size_t)i] = 0;
        }
        for (Index i = n - 1; i >= 0; --i)
            for (Index j = i + 1; j < n; ++j) y[(size_t)i] -= m_.coeff(j, i) * y[(size_t)j];
        for (Index k = n - 1; k >= 0; --k) if (tr_[(size_t)k] != k) std::swap(y[(size_t)k], y[(size_t)tr_[(size_t)k]]);
        Matrix<Scalar, M::RowsAtCompileTime, 1> x;
        x.resizeLike(n, 1);
        for (Index i = 0; i < n; ++i) x.coeffRef(i, 0) = y[(size_t)i];
        return x;
    }
private:
    M m_;
    Index n_;
    std::vector<Index> tr_;
};
template <class D> LDLT<typename Base<D>::Plain> Base<D>::ldlt() const { return LDLT<Plain>(eval()); }

// ------------------------------------------------------------------------------------------------------------------------
// the typedefs the reference uses
// ------------------------------------------------------------------------------------------------------------------------
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<int, Dynamic, Dynamic> MatrixXi;
typedef Matrix<double, Dynamic, 1> VectorXd;

}  // namespace Eigen

#endif
