// Minimal stand-in for the slice of Boost.uBLAS the reference touches (dense (i,j) container only).
// ORACLE BUILD ONLY: lets the unmodified reference headers compile in a container without Boost.
#ifndef VA_SHIM_UBLAS_MATRIX_HPP
#define VA_SHIM_UBLAS_MATRIX_HPP
#include <cstddef>
#include <vector>
namespace boost { namespace numeric { namespace ublas {
template <class T>
class matrix
{
    std::size_t r_ = 0, c_ = 0;
    std::vector<T> d_;

  public:
    matrix() = default;
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c) {}
    T &operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    const T &operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
};
}}} // namespace boost::numeric::ublas
#endif
