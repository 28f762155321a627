// Compiles against include/tokenmonster_gpu.hpp and links libtmgpu.so: the C++ mirror of the
// reference's tokenmonster-cpp API.  Without a GPU it checks the error path (no CPU fallback);
// with a GPU (argv[2] == "gpu") it runs the reference's tests/unit.cpp:87-112 assertions.
#include <cassert>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>

#include "tokenmonster_gpu.hpp"

static std::span<const std::uint8_t> bytes(std::string_view s) {
  return {reinterpret_cast<const std::uint8_t*>(s.data()), s.size()};
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const bool gpu = argc > 2 && std::strcmp(argv[2], "gpu") == 0;
  if (!gpu) {
    try {
      auto v = tokenmonster_gpu::Vocab::load(argv[1], 0);
      (void)v;
      if (tmgpu_device_count() == 0) return 1;  // must not succeed without a device
    } catch (const tokenmonster_gpu::Error& e) {
      std::printf("expected: %d %s\n", e.code(), e.what());
      return e.code() == TMGPU_E_CUDA ? 0 : 1;
    }
    return 0;
  }
  auto vocab = tokenmonster_gpu::Vocab::load(argv[1], 0);
  assert(vocab.size() == 4 && vocab.max_token_length() == 2);
  auto result = vocab.tokenize(bytes("ab a z"));
  assert((result.tokens == std::vector<std::uint32_t>{3, 0, 1, 0}) && result.missing == 1);
  auto decoded = vocab.decode(result.tokens);
  assert(std::string(decoded.begin(), decoded.end()) == "ab a ");
  auto ser = vocab.tokenize_serialized(bytes("ab"), 2);
  assert(ser.encoding_length == 2 && (ser.bytes == std::vector<std::uint8_t>{3, 0}));
  auto cnt = vocab.count(bytes("ab a z"));
  assert(cnt.tokens == 4 && cnt.missing == 1);
  auto dec = vocab.new_decoder();
  auto s1 = dec.decode({{3, 0, 1}});
  assert(std::string(s1[0].begin(), s1[0].end()) == "ab a");
  assert(dec.flush()[0].empty());
  std::puts("cpp unit.cpp assertions: PASS");
  return 0;
}
