// hpxfft::util::vector_2d<T> -- drop-in for core/include/hpxfft/util/vector_2d.hpp of HPX-FFT.
//
// Same public surface (fields values_, size_, n_row_, n_col_; operator()(i,j) = values_[i*n_col_+j];
// row(i); bounds-checked at() throwing std::runtime_error; exact operator==; begin/end/data/size),
// so code written against the reference compiles unchanged.  Differences, all deliberate:
//   * storage is page-locked (hpxfft_b200_host_alloc) when the CUDA library can provide it, so that the
//     by-value hand-over to loop::initialize() runs at PCIe speed; falls back to new[] otherwise;
//   * the destructor frees (the reference's is defaulted and leaks, vector_2d.hpp:31);
//   * copy assignment copies (the reference's lvalue operator= swaps, vector_2d.hpp:146-151).
// HPX serialisation is compiled in only when HPX headers are present.
#ifndef HPXFFT_B200_VECTOR_2D_HPP
#define HPXFFT_B200_VECTOR_2D_HPP

#include "../../hpxfft_b200.h"

#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <utility>

#if defined(HPXFFT_B200_WITH_HPX)
#include <hpx/serialization.hpp>
#endif

namespace hpxfft::util
{
template <typename T>
struct vector_2d
{
    T *values_;
    std::size_t size_;
    std::size_t n_row_;  // first dimension (row-major)
    std::size_t n_col_;  // second dimension

    using iterator = T *;
    using const_iterator = const T *;

    vector_2d() : values_(nullptr), size_(0), n_row_(0), n_col_(0), pinned_(false) {}
    vector_2d(std::size_t n_row, std::size_t n_col) : vector_2d(n_row, n_col, T()) {}
    vector_2d(std::size_t n_row, std::size_t n_col, const T &v) : vector_2d()
    {
        allocate(n_row, n_col);
        std::fill(begin(), end(), v);
    }
    vector_2d(const vector_2d &src) : vector_2d()
    {
        allocate(src.n_row_, src.n_col_);
        std::copy(src.begin(), src.end(), begin());
    }
    vector_2d(vector_2d &&mv) noexcept : vector_2d() { swap(*this, mv); }
    ~vector_2d() { release(); }

    vector_2d &operator=(const vector_2d &src)
    {
        if (this != &src)
        {
            vector_2d tmp(src);
            swap(*this, tmp);
        }
        return *this;
    }
    vector_2d &operator=(vector_2d &&mv) noexcept
    {
        swap(*this, mv);
        return *this;
    }

    T &operator()(std::size_t i, std::size_t j) { return values_[i * n_col_ + j]; }
    const T &operator()(std::size_t i, std::size_t j) const { return values_[i * n_col_ + j]; }
    T &at(std::size_t i, std::size_t j)
    {
        if (i * n_col_ + j >= size_) throw std::runtime_error("out of range exception");
        return values_[i * n_col_ + j];
    }
    const T &at(std::size_t i, std::size_t j) const
    {
        if (i * n_col_ + j >= size_) throw std::runtime_error("out of range exception");
        return values_[i * n_col_ + j];
    }
    constexpr T *data() noexcept { return values_; }
    constexpr const T *data() const noexcept { return values_; }

    iterator begin() noexcept { return values_; }
    const_iterator begin() const noexcept { return values_; }
    iterator end() noexcept { return values_ + size_; }
    const_iterator end() const noexcept { return values_ + size_; }
    const_iterator cbegin() const noexcept { return values_; }
    const_iterator cend() const noexcept { return values_ + size_; }
    iterator row(std::size_t i) noexcept { return values_ + i * n_col_; }
    const_iterator row(std::size_t i) const noexcept { return values_ + i * n_col_; }

    std::size_t size() const noexcept { return size_; }
    std::size_t n_row() const noexcept { return n_row_; }
    std::size_t n_col() const noexcept { return n_col_; }

    friend void swap(vector_2d &a, vector_2d &b) noexcept
    {
        std::swap(a.values_, b.values_);
        std::swap(a.size_, b.size_);
        std::swap(a.n_row_, b.n_row_);
        std::swap(a.n_col_, b.n_col_);
        std::swap(a.pinned_, b.pinned_);
    }

  private:
    bool pinned_;

    void allocate(std::size_t n_row, std::size_t n_col)
    {
        n_row_ = n_row;
        n_col_ = n_col;
        size_ = n_row * n_col;
        if (size_ == 0) return;
        void *p = hpxfft_b200_device_count() > 0 ? hpxfft_b200_host_alloc(size_ * sizeof(T)) : nullptr;
        pinned_ = p != nullptr;
        values_ = pinned_ ? static_cast<T *>(p) : new T[size_];
    }
    void release() noexcept
    {
        if (values_)
        {
            if (pinned_)
                hpxfft_b200_host_free(values_);
            else
                delete[] values_;
        }
        values_ = nullptr;
        size_ = n_row_ = n_col_ = 0;
    }

#if defined(HPXFFT_B200_WITH_HPX)
    friend class hpx::serialization::access;
    template <typename Archive>
    void save(Archive &ar, const unsigned int) const
    {
        ar << n_row_ << n_col_ << size_;
        for (std::size_t i = 0; i < size_; ++i) ar << values_[i];
    }
    template <typename Archive>
    void load(Archive &ar, const unsigned int)
    {
        std::size_t r, c, s;
        ar >> r >> c >> s;
        release();
        allocate(r, c);
        for (std::size_t i = 0; i < size_; ++i) ar >> values_[i];
    }
    HPX_SERIALIZATION_SPLIT_MEMBER()
#endif
};

template <typename H>
inline bool operator==(const vector_2d<H> &lhs, const vector_2d<H> &rhs)
{
    if (lhs.n_row_ != rhs.n_row_ || lhs.n_col_ != rhs.n_col_) return false;
    return std::equal(lhs.begin(), lhs.end(), rhs.begin());
}
}  // namespace hpxfft::util
#endif
