// Stand-in for the third-party header stl_reader.h (not in the reference tree) so that BaseLib/motion/Mesh.cpp compiles.
// OBJloader::Mesh::loadSTL is never reached by the pinned paths; constructing a StlMesh aborts.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <sstream>
#include <stdexcept>
#define STL_READER_COND_THROW(cond, msg) if (cond) { std::stringstream ss_; ss_ << msg; throw std::runtime_error(ss_.str()); }
namespace stl_reader {
template <class TNumber, class TIndex> class StlMesh {
 public:
  explicit StlMesh(const char*) { std::fprintf(stderr, "oracle/_ref: stl_reader is a stub\n"); std::abort(); }
  size_t num_vrts() const { return 0; }
  size_t num_tris() const { return 0; }
  const TNumber* vrt_coords(size_t) const { return nullptr; }
  TIndex tri_corner_ind(size_t, size_t) const { return 0; }
  const TNumber* tri_normal(size_t) const { return nullptr; }
};
}  // namespace stl_reader
