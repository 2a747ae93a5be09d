// strumpack_b200 -- HSS construction on the GPU (placeholder, see header).
#include "hss_compress.hpp"
#include <stdexcept>
namespace sb200 {
HSSHost compress_dense(int, int, const double*, int, const CompressOptions&) {
  throw std::runtime_error("compress_dense: not implemented yet");
}
HSSHost compress_elements(int, int, double (*)(int, int), const CompressOptions&) {
  throw std::runtime_error("compress_elements: not implemented yet");
}
HSSHost compress_kernel(int, int, double*, int, double, double,
                        const CompressOptions&, int*) {
  throw std::runtime_error("compress_kernel: not implemented yet");
}
}  // namespace sb200
