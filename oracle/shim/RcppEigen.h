// Stand-in for <RcppEigen.h> / <Rcpp.h> so that the reference's OWN model sources
// (/root/reference/pkg/src/{model,gap_factor_model,zi_gap_factor_model,sparse_gap_factor_model,
// zi_sparse_gap_factor_model}.cpp, algorithm*.h, utils/*.h) compile unmodified into oracle/_ref/libpcmf_ref.so
// on a machine that has neither R, Rcpp, Eigen nor Boost.  TEST INFRASTRUCTURE ONLY: it exists to pin the CPU oracle
// (oracle/pcmf_oracle.c) against the reference's real control flow and formulas; nothing under pcmf_b200/ uses it.
//
// This is NOT Eigen.  It is a small eager dense-matrix class written for this purpose, covering exactly the
// expression forms those files use (listed in oracle/README_ref.md).  Every operation evaluates immediately into a
// new column-major matrix; `.array()` only flips the meaning of `*` and `/` to element-wise, as in Eigen.
// Differences that matter and are documented there: summation order (plain loops), vectors carry no static
// orientation (a size-matched vector is accepted wherever Eigen would transpose implicitly), and (i, j) access to a
// vector ignores the out-of-range index -- the reference's bernoulli_loglike_vec reads prob(i, j) on a 1 x cols
// vector (likelihood.h:307), which is undefined behaviour in the real build.
#ifndef PCMF_REF_SHIM_RCPPEIGEN_H
#define PCMF_REF_SHIM_RCPPEIGEN_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <type_traits>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
inline void initParallel() {}

template <typename T> class Dense;
template <typename T> class Arr;
template <typename T> class Block;
template <typename T, bool A> class Vecwise;

class PermutationMatrixX {
public:
    explicit PermutationMatrixX(int n = 0) : m_idx(n) { for (int i = 0; i < n; ++i) m_idx[i] = i; }
    std::vector<int> &indices() { return m_idx; }
    const std::vector<int> &indices() const { return m_idx; }
private:
    std::vector<int> m_idx;
};
template <int A, int B> using PermutationMatrix = PermutationMatrixX;

// "Matrix": `*` is the matrix product
template <typename T>
class Dense {
public:
    typedef T Scalar;
    Dense() : m_r(0), m_c(0) {}
    explicit Dense(int n) : m_r(n), m_c(1), m_d((size_t)n) {}
    Dense(int r, int c) : m_r(r), m_c(c), m_d((size_t)r * c) {}
    Dense(const Block<T> &b);

    static Dense Constant(int r, int c, T v) { Dense m(r, c); std::fill(m.m_d.begin(), m.m_d.end(), v); return m; }
    static Dense Ones(int r, int c) { return Constant(r, c, T(1)); }
    static Dense Zero(int r, int c) { return Constant(r, c, T(0)); }
    static Dense Ones(int n) { return Constant(n, 1, T(1)); }
    static Dense Zero(int n) { return Constant(n, 1, T(0)); }
    static Dense LinSpaced(int n, T lo, T hi) { Dense m(n); for (int i = 0; i < n; ++i) m[i] = n > 1 ? (T)(lo + (hi - lo) * i / (n - 1)) : hi; return m; }

    int rows() const { return m_r; }
    int cols() const { return m_c; }
    int size() const { return m_r * m_c; }
    void resize(int r, int c) { m_r = r; m_c = c; m_d.assign((size_t)r * c, T()); }
    T *data() { return m_d.data(); }
    const T *data() const { return m_d.data(); }

    // element access; on a vector the out-of-range index of an (i, j) pair is ignored (see the header comment)
    T &operator()(int i, int j) { return m_d[idx(i, j)]; }
    const T &operator()(int i, int j) const { return m_d[idx(i, j)]; }
    T &operator()(int i) { return m_d[(size_t)i]; }
    const T &operator()(int i) const { return m_d[(size_t)i]; }
    T &operator[](int i) { return m_d[(size_t)i]; }
    const T &operator[](int i) const { return m_d[(size_t)i]; }

    Arr<T> array() const;
    Dense matrix() const { return *this; }
    Dense eval() const { return *this; }
    Dense transpose() const {
        Dense m(m_c, m_r);
        for (int j = 0; j < m_c; ++j) for (int i = 0; i < m_r; ++i) m.m_d[(size_t)j + (size_t)i * m_c] = m_d[(size_t)i + (size_t)j * m_r];
        return m;
    }
    template <typename U> Dense<U> cast() const {
        Dense<U> m(m_r, m_c);
        for (size_t e = 0; e < m_d.size(); ++e) m.data()[e] = (U)m_d[e];
        return m;
    }
    template <typename F> Dense unaryExpr(F f) const {
        Dense m(m_r, m_c);
        for (size_t e = 0; e < m_d.size(); ++e) m.m_d[e] = f(m_d[e]);
        return m;
    }
    T sum() const { T s = T(0); for (size_t e = 0; e < m_d.size(); ++e) s += m_d[e]; return s; }
    T mean() const { return sum() / (T)size(); }
    T maxCoeff() const { T s = m_d[0]; for (size_t e = 1; e < m_d.size(); ++e) if (m_d[e] > s) s = m_d[e]; return s; }
    T minCoeff() const { T s = m_d[0]; for (size_t e = 1; e < m_d.size(); ++e) if (m_d[e] < s) s = m_d[e]; return s; }
    T squaredNorm() const { T s = T(0); for (size_t e = 0; e < m_d.size(); ++e) s += m_d[e] * m_d[e]; return s; }
    T trace() const { T s = T(0); for (int i = 0; i < std::min(m_r, m_c); ++i) s += (*this)(i, i); return s; }
    Dense cwiseMax(T v) const { Dense m(*this); for (size_t e = 0; e < m_d.size(); ++e) m.m_d[e] = std::max(m_d[e], v); return m; }
    Dense cwiseProduct(const Dense &o) const { Dense m(*this); check_size(o); for (size_t e = 0; e < m_d.size(); ++e) m.m_d[e] = m_d[e] * o.m_d[e]; return m; }
    Dense leftCols(int k) const { Dense m(m_r, k); std::copy(m_d.begin(), m_d.begin() + (size_t)m_r * k, m.m_d.begin()); return m; }
    Dense head(int k) const { Dense m(k); std::copy(m_d.begin(), m_d.begin() + k, m.m_d.begin()); return m; }

    Block<T> row(int i) { return Block<T>(this, i, true); }
    Block<T> col(int j) { return Block<T>(this, j, false); }
    Block<T> row(int i) const { return Block<T>(const_cast<Dense *>(this), i, true); }
    Block<T> col(int j) const { return Block<T>(const_cast<Dense *>(this), j, false); }
    Vecwise<T, false> rowwise() const;
    Vecwise<T, false> colwise() const;

    Dense &operator+=(const Dense &o) { check_size(o); for (size_t e = 0; e < m_d.size(); ++e) m_d[e] += o.m_d[e]; return *this; }
    Dense &operator-=(const Dense &o) { check_size(o); for (size_t e = 0; e < m_d.size(); ++e) m_d[e] -= o.m_d[e]; return *this; }
    Dense &operator*=(const PermutationMatrixX &p) { *this = permuted(p); return *this; }
    // (M * P)(:, c) = M(:, indices[c])   (P has a 1 at (indices[c], c))
    Dense permuted(const PermutationMatrixX &p) const {
        Dense m(m_r, m_c);
        for (int c = 0; c < m_c; ++c) for (int i = 0; i < m_r; ++i) m(i, c) = (*this)(i, p.indices()[c]);
        return m;
    }
    void check_size(const Dense &o) const {
        if (o.size() != size()) throw std::runtime_error("shim: size mismatch in element-wise operation");
    }
private:
    size_t idx(int i, int j) const {
        if (m_r == 1) return (size_t)j;
        if (m_c == 1) return (size_t)i;
        return (size_t)i + (size_t)j * m_r;
    }
    int m_r, m_c;
    std::vector<T> m_d;
};

// "Array": same storage, `*` and `/` are element-wise; converts back to a matrix on assignment (slicing)
template <typename T>
class Arr : public Dense<T> {
public:
    Arr() {}
    Arr(const Dense<T> &m) : Dense<T>(m) {}
    Arr array() const { return *this; }
    Dense<T> matrix() const { return Dense<T>(*this); }
    template <typename U> Arr<U> cast() const { return Arr<U>(Dense<T>::template cast<U>()); }
    template <typename F> Arr unaryExpr(F f) const { return Arr(Dense<T>::unaryExpr(f)); }
    Vecwise<T, true> rowwise() const;
    Vecwise<T, true> colwise() const;
};
template <typename T> Arr<T> Dense<T>::array() const { return Arr<T>(*this); }

// one row or one column of a matrix: readable as a vector, assignable from any vector of the same length
template <typename T>
class Block {
public:
    Block(Dense<T> *m, int k, bool is_row) : m_m(m), m_k(k), m_row(is_row) {}
    int size() const { return m_row ? m_m->cols() : m_m->rows(); }
    T get(int e) const { return m_row ? (*m_m)(m_k, e) : (*m_m)(e, m_k); }
    Dense<T> eval() const {
        Dense<T> v = m_row ? Dense<T>(1, size()) : Dense<T>(size(), 1);
        for (int e = 0; e < size(); ++e) v[e] = get(e);
        return v;
    }
    Block &operator=(const Dense<T> &v) {
        if (v.size() != size()) throw std::runtime_error("shim: block assignment size mismatch");
        for (int e = 0; e < size(); ++e) { if (m_row) (*m_m)(m_k, e) = v[e]; else (*m_m)(e, m_k) = v[e]; }
        return *this;
    }
    Block &operator=(const Block &o) { return (*this = o.eval()); }
    Dense<T> transpose() const { return eval().transpose(); }
    T sum() const { return eval().sum(); }
    T mean() const { return eval().mean(); }
    template <typename U> Dense<U> cast() const { return eval().template cast<U>(); }
    Arr<T> array() const { return eval().array(); }
private:
    Dense<T> *m_m; int m_k; bool m_row;
};
template <typename T> Dense<T>::Dense(const Block<T> &b) { *this = b.eval(); }

// rowwise() / colwise(): reductions, and broadcasting of a vector over the rows / columns
template <typename T, bool A>
class Vecwise {
public:
    typedef typename std::conditional<A, Arr<T>, Dense<T> >::type Out;
    Vecwise(const Dense<T> &m, bool rowwise) : m_m(m), m_row(rowwise) {}
    Dense<T> sum() const {
        if (m_row) { Dense<T> v = Dense<T>::Zero(m_m.rows(), 1); for (int j = 0; j < m_m.cols(); ++j) for (int i = 0; i < m_m.rows(); ++i) v[i] += m_m(i, j); return v; }
        Dense<T> v = Dense<T>::Zero(1, m_m.cols()); for (int j = 0; j < m_m.cols(); ++j) for (int i = 0; i < m_m.rows(); ++i) v[j] += m_m(i, j); return v;
    }
    Dense<T> mean() const {
        Dense<T> v = sum(); const T d = (T)(m_row ? m_m.cols() : m_m.rows());
        for (int e = 0; e < v.size(); ++e) v[e] /= d;
        return v;
    }
    template <typename F> Out bcast(const Dense<T> &v, F f) const {
        if (v.size() != (m_row ? m_m.cols() : m_m.rows())) throw std::runtime_error("shim: broadcast size mismatch");
        Dense<T> r(m_m.rows(), m_m.cols());
        for (int j = 0; j < m_m.cols(); ++j) for (int i = 0; i < m_m.rows(); ++i) r(i, j) = f(m_m(i, j), v[m_row ? j : i]);
        return Out(r);
    }
    Out operator+(const Dense<T> &v) const { return bcast(v, [](T a, T b) { return a + b; }); }
    Out operator-(const Dense<T> &v) const { return bcast(v, [](T a, T b) { return a - b; }); }
    Out operator*(const Dense<T> &v) const { return bcast(v, [](T a, T b) { return a * b; }); }
    Out operator/(const Dense<T> &v) const { return bcast(v, [](T a, T b) { return a / b; }); }
private:
    Dense<T> m_m; bool m_row;
};
template <typename T> Vecwise<T, false> Dense<T>::rowwise() const { return Vecwise<T, false>(*this, true); }
template <typename T> Vecwise<T, false> Dense<T>::colwise() const { return Vecwise<T, false>(*this, false); }
template <typename T> Vecwise<T, true> Arr<T>::rowwise() const { return Vecwise<T, true>(*this, true); }
template <typename T> Vecwise<T, true> Arr<T>::colwise() const { return Vecwise<T, true>(*this, false); }

// ---- binary operators ---------------------------------------------------------------------------------------
template <typename T, typename F> Dense<T> ewise(const Dense<T> &a, const Dense<T> &b, F f) {
    a.check_size(b);
    Dense<T> r(a.rows(), a.cols());
    for (int e = 0; e < a.size(); ++e) r[e] = f(a[e], b[e]);
    return r;
}
template <typename T> Dense<T> operator+(const Dense<T> &a, const Dense<T> &b) { return ewise(a, b, [](T x, T y) { return x + y; }); }
template <typename T> Dense<T> operator-(const Dense<T> &a, const Dense<T> &b) { return ewise(a, b, [](T x, T y) { return x - y; }); }
template <typename T> Dense<T> operator*(const Dense<T> &a, const Dense<T> &b) {
    if (a.cols() != b.rows()) throw std::runtime_error("shim: matrix product size mismatch");
    Dense<T> r = Dense<T>::Zero(a.rows(), b.cols());
    for (int j = 0; j < b.cols(); ++j)
        for (int k = 0; k < a.cols(); ++k) {
            const T bkj = b(k, j);
            for (int i = 0; i < a.rows(); ++i) r.data()[(size_t)i + (size_t)j * a.rows()] += a.data()[(size_t)i + (size_t)k * a.rows()] * bkj;
        }
    return r;
}
template <typename T> Arr<T> operator+(const Arr<T> &a, const Arr<T> &b) { return Arr<T>(ewise<T>(a, b, [](T x, T y) { return x + y; })); }
template <typename T> Arr<T> operator-(const Arr<T> &a, const Arr<T> &b) { return Arr<T>(ewise<T>(a, b, [](T x, T y) { return x - y; })); }
template <typename T> Arr<T> operator*(const Arr<T> &a, const Arr<T> &b) { return Arr<T>(ewise<T>(a, b, [](T x, T y) { return x * y; })); }
template <typename T> Arr<T> operator/(const Arr<T> &a, const Arr<T> &b) { return Arr<T>(ewise<T>(a, b, [](T x, T y) { return x / y; })); }
template <typename T> Dense<T> operator*(const Dense<T> &a, const PermutationMatrixX &p) { return a.permuted(p); }
template <typename T> Dense<T> operator+(const Block<T> &a, const Block<T> &b) { return a.eval() + b.eval(); }
template <typename T> Dense<T> operator-(const Block<T> &a, const Block<T> &b) { return a.eval() - b.eval(); }
#define PCMF_SHIM_SCALAR_OPS(OP)                                                                                                  \
    template <typename T> Dense<T> operator OP(const Dense<T> &a, double s) { Dense<T> r(a); for (int e = 0; e < r.size(); ++e) r[e] = (T)(a[e] OP s); return r; } \
    template <typename T> Dense<T> operator OP(double s, const Dense<T> &a) { Dense<T> r(a); for (int e = 0; e < r.size(); ++e) r[e] = (T)(s OP a[e]); return r; } \
    template <typename T> Arr<T> operator OP(const Arr<T> &a, double s) { Arr<T> r(a); for (int e = 0; e < r.size(); ++e) r[e] = (T)(a[e] OP s); return r; } \
    template <typename T> Arr<T> operator OP(double s, const Arr<T> &a) { Arr<T> r(a); for (int e = 0; e < r.size(); ++e) r[e] = (T)(s OP a[e]); return r; }
PCMF_SHIM_SCALAR_OPS(+)
PCMF_SHIM_SCALAR_OPS(-)
PCMF_SHIM_SCALAR_OPS(*)
PCMF_SHIM_SCALAR_OPS(/)
#undef PCMF_SHIM_SCALAR_OPS
template <typename T> Dense<T> operator-(const Dense<T> &a) { Dense<T> r(a); for (int e = 0; e < r.size(); ++e) r[e] = -a[e]; return r; }
template <typename T> Arr<unsigned char> operator>(const Arr<T> &a, double s) { Dense<unsigned char> r(a.rows(), a.cols()); for (int e = 0; e < a.size(); ++e) r[e] = a[e] > s; return Arr<unsigned char>(r); }
template <typename T> Arr<unsigned char> operator>=(const Arr<T> &a, double s) { Dense<unsigned char> r(a.rows(), a.cols()); for (int e = 0; e < a.size(); ++e) r[e] = a[e] >= s; return Arr<unsigned char>(r); }
template <typename T> std::ostream &operator<<(std::ostream &o, const Dense<T> &m) {
    for (int i = 0; i < m.rows(); ++i) { for (int j = 0; j < m.cols(); ++j) o << m(i, j) << " "; o << "\n"; }
    return o;
}

typedef Dense<double> MatrixXd;
typedef Dense<double> VectorXd;
typedef Dense<double> RowVectorXd;
typedef Arr<double> ArrayXXd;
typedef Dense<int> MatrixXi;
typedef Dense<int> VectorXi;
template <typename M> using Map = M;

}  // namespace Eigen

// ---- Rcpp ---------------------------------------------------------------------------------------------------------
typedef void *SEXP;
namespace Rcpp {

static std::ostream &Rcout = std::cout;
inline void checkUserInterrupt() {}
[[noreturn]] inline void stop(const std::string &msg) { throw std::runtime_error(msg); }

class List;
struct Any {
    enum Kind { NONE, MATD, MATI, DBL, INT, BOOL, LIST } kind = NONE;
    Eigen::Dense<double> md; Eigen::Dense<int> mi; double d = 0; int i = 0; bool b = false; std::shared_ptr<List> l;
    Any() {}
    Any(const Eigen::Dense<double> &m) : kind(MATD), md(m) {}
    Any(const Eigen::Dense<int> &m) : kind(MATI), mi(m) {}
    Any(double v) : kind(DBL), d(v) {}
    Any(int v) : kind(INT), i(v) {}
    Any(bool v) : kind(BOOL), b(v) {}
    Any(const List &v);
};
struct NamedValue { std::string name; Any value; };
struct Named {
    std::string name;
    explicit Named(const std::string &n) : name(n) {}
    template <typename T> NamedValue operator=(const T &v) const { return NamedValue{name, Any(v)}; }
};
class List {
public:
    std::vector<NamedValue> items;
    static List create() { return List(); }
    template <typename... Args> static List create(const Args &...args) { List l; int dummy[] = {0, (l.items.push_back(args), 0)...}; (void)dummy; return l; }
    template <typename T> void push_back(const T &v, const std::string &name) { items.push_back(NamedValue{name, Any(v)}); }
    const Any *find(const std::string &name) const { for (const auto &it : items) if (it.name == name) return &it.value; return nullptr; }
};
inline Any::Any(const List &v) : kind(LIST), l(std::make_shared<List>(v)) {}

}  // namespace Rcpp

#endif
