// Stand-in for boost::multi_array<T,2>: value-initialised (zero) storage like the real one,
// extents[a][b] constructor and arr[i][j] access. Nothing else is used by the reference.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
  struct shim_extent2 { std::size_t a, b; };
  struct shim_extent1 {
    std::size_t a;
    shim_extent2 operator[](std::size_t b) const { return shim_extent2{a, b}; }
  };
  struct shim_extent0 {
    shim_extent1 operator[](std::size_t a) const { return shim_extent1{a}; }
  };
  static const shim_extent0 extents = {};
  template <class T, std::size_t D> class multi_array;
  template <class T> class multi_array<T, 2> {
  public:
    explicit multi_array(const shim_extent2& e) : _cols(e.b), _data(e.a * e.b) {}
    T* operator[](std::size_t i) { return _data.data() + i * _cols; }
    const T* operator[](std::size_t i) const { return _data.data() + i * _cols; }
  private:
    std::size_t _cols;
    std::vector<T> _data;  // value-initialised => zeros
  };
}
