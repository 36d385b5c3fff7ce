// hpx_stub (see README.md): byte-vector archives with the operator<< / operator>> surface and the split-member macro.  NOT HPX.
#pragma once
#include <cstddef>
#include <cstring>
#include <type_traits>
#include <vector>
namespace hpx::serialization
{
// the friend through which archives reach a type's private serialize / save / load
class access
{
  public:
    template <class Archive, class T> static void serialize(Archive &ar, T &v, unsigned version) { v.serialize(ar, version); }
};
struct output_archive
{
    static constexpr bool is_saving = true;
    std::vector<char> &buf;
    explicit output_archive(std::vector<char> &b) : buf(b) {}
    template <class T> std::enable_if_t<std::is_arithmetic_v<T>, output_archive &> operator<<(T const &v)
    {
        const char *p = reinterpret_cast<const char *>(&v);
        buf.insert(buf.end(), p, p + sizeof(T));
        return *this;
    }
    template <class T> std::enable_if_t<!std::is_arithmetic_v<T>, output_archive &> operator<<(T const &v)
    {
        access::serialize(*this, const_cast<T &>(v), 0u);
        return *this;
    }
};
struct input_archive
{
    static constexpr bool is_saving = false;
    std::vector<char> const &buf;
    std::size_t pos = 0;
    explicit input_archive(std::vector<char> const &b) : buf(b) {}
    template <class T> std::enable_if_t<std::is_arithmetic_v<T>, input_archive &> operator>>(T &v)
    {
        std::memcpy(&v, buf.data() + pos, sizeof(T));
        pos += sizeof(T);
        return *this;
    }
    template <class T> std::enable_if_t<!std::is_arithmetic_v<T>, input_archive &> operator>>(T &v)
    {
        access::serialize(*this, v, 0u);
        return *this;
    }
};
}  // namespace hpx::serialization
#define HPX_SERIALIZATION_SPLIT_MEMBER()                                                                     \
    template <typename Archive> void serialize(Archive &ar, const unsigned int v)                          \
    {                                                                                                      \
        if constexpr (Archive::is_saving)                                                                  \
            save(ar, v);                                                                                   \
        else                                                                                               \
            load(ar, v);                                                                                   \
    }
