// conversion.hpp — the byte layouts SemaDB persists under a shard bucket, restated for the C++
// host side of the GPU index (conversion/conversion.go:57-124, conversion/keys.go:6-20).
// Byte strings are std::string (may hold NULs), like Go's []byte used as map keys.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace semadb {
namespace conversion {

// conversion.go:57-65 — little endian u64
inline std::string Uint64ToBytes(uint64_t v) {
  std::string b(8, '\0');
  for (int i = 0; i < 8; ++i) b[i] = char((v >> (8 * i)) & 0xFF);
  return b;
}
inline uint64_t BytesToUint64(const std::string& b) {
  uint64_t v = 0;
  for (int i = 0; i < 8 && i < int(b.size()); ++i) v |= uint64_t(uint8_t(b[i])) << (8 * i);
  return v;
}

// conversion.go:78-104 — IEEE-754 bits, little endian, 4 bytes per element
inline std::string Float32ToBytes(const float* f, size_t n) {
  std::string b(n * 4, '\0');
  for (size_t i = 0; i < n; ++i) {
    uint32_t u;
    std::memcpy(&u, f + i, 4);
    for (int k = 0; k < 4; ++k) b[i * 4 + k] = char((u >> (8 * k)) & 0xFF);
  }
  return b;
}
inline std::string Float32ToBytes(const std::vector<float>& f) { return Float32ToBytes(f.data(), f.size()); }
inline std::vector<float> BytesToFloat32(const std::string& b) {
  std::vector<float> f(b.size() / 4);
  for (size_t i = 0; i < f.size(); ++i) {
    uint32_t u = 0;
    for (int k = 0; k < 4; ++k) u |= uint32_t(uint8_t(b[i * 4 + k])) << (8 * k);
    std::memcpy(&f[i], &u, 4);
  }
  return f;
}

// conversion.go:110-124 — edge lists (and bit-packed vectors, binary.go:283) as LE u64
inline std::string EdgeListToBytes(const uint64_t* e, size_t n) {
  std::string b(n * 8, '\0');
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 8; ++k) b[i * 8 + k] = char((e[i] >> (8 * k)) & 0xFF);
  return b;
}
inline std::string EdgeListToBytes(const std::vector<uint64_t>& e) { return EdgeListToBytes(e.data(), e.size()); }
inline std::vector<uint64_t> BytesToEdgeList(const std::string& b) {
  std::vector<uint64_t> e(b.size() / 8);
  for (size_t i = 0; i < e.size(); ++i) {
    uint64_t v = 0;
    for (int k = 0; k < 8; ++k) v |= uint64_t(uint8_t(b[i * 8 + k])) << (8 * k);
    e[i] = v;
  }
  return e;
}

// keys.go:6-12 — 'n' + u64le(id) + suffix ('v' vector, 'e' edges, 'q' quantized)
inline std::string NodeKey(uint64_t id, char suffix) {
  std::string k(10, '\0');
  k[0] = 'n';
  for (int i = 0; i < 8; ++i) k[1 + i] = char((id >> (8 * i)) & 0xFF);
  k[9] = suffix;
  return k;
}
// keys.go:15-20
inline bool NodeIdFromKey(const std::string& key, char suffix, uint64_t* id) {
  if (key.size() != 10 || key[0] != 'n' || key[9] != suffix) return false;
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v |= uint64_t(uint8_t(key[1 + i])) << (8 * i);
  *id = v;
  return true;
}

}  // namespace conversion
}  // namespace semadb
