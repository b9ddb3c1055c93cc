// The raw/<prefix>_<step>.zst container of include/SubrosaDG_b200/SubrosaDG.hpp (RawBinaryCompress, after src/View/RawBinary.cpp:42-74):
// every payload is written with libzstd (as the reference does) and with the self-contained raw-block frame writer, and every file is
// read back with both readers.  usage: raw_binary_container DIR SIZE...   (payload = SIZE bytes of a fixed pattern; files DIR/{lib,raw}_SIZE.zst)
//                                        raw_binary_container decode FILE OUT   (the payload of a file written elsewhere, e.g. by the reference)
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <fstream>
#include <iostream>

static std::string pattern(std::size_t n) {
  std::string s(n, '\0');
  std::uint64_t x = 0x9E3779B97F4A7C15ULL;
  for (std::size_t i = 0; i < n; i++) {
    if (i % 4096 < 1024) { s[i] = static_cast<char>(i / 4096); continue; }   // compressible stretches
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    s[i] = static_cast<char>(x & 0xFF);
  }
  return s;
}

int main(int argc, char* argv[]) {
  using SubrosaDG::RawBinaryCompress;
  if (argc < 3) return 2;
  if (std::string(argv[1]) == "decode") {   // raw_binary_container decode FILE OUT: RawBinaryCompress::read of a file someone else wrote
    if (argc < 4) return 2;
    std::stringstream back;
    RawBinaryCompress::read(argv[2], back);
    const std::string payload = back.str();
    std::ofstream(argv[3], std::ios::binary).write(payload.data(), static_cast<std::streamsize>(payload.size()));
    std::cout << "decoded " << payload.size() << " bytes\n";
    return 0;
  }
  const std::filesystem::path dir(argv[1]);
  const bool have_lib = RawBinaryCompress::zstd().ok();
  std::cout << "libzstd " << (have_lib ? "found" : "absent") << "\n";
  for (int a = 2; a < argc; a++) {
    const std::size_t n = static_cast<std::size_t>(std::atoll(argv[a]));
    const std::string payload = pattern(n);
    for (int writer = 0; writer < 2; writer++) {
      if (writer == 0 && !have_lib) continue;
      RawBinaryCompress::use_system_zstd = writer == 0;
      std::stringstream ss;
      ss.write(payload.data(), static_cast<std::streamsize>(n));
      const std::filesystem::path path = dir / ((writer == 0 ? "lib_" : "raw_") + std::to_string(n) + ".zst");
      RawBinaryCompress::write(path, ss);
      for (int reader = 0; reader < 2; reader++) {
        if (reader == 0 && !have_lib) continue;
        if (reader == 1 && writer == 0) continue;   // compressed blocks need the library
        RawBinaryCompress::use_system_zstd = reader == 0;
        std::stringstream back;
        RawBinaryCompress::read(path, back);
        if (back.str() != payload) { std::cout << "MISMATCH size " << n << " writer " << writer << " reader " << reader << "\n"; return 1; }
      }
    }
  }
  std::cout << "OK\n";
  return 0;
}
