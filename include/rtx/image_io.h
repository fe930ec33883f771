// image_io.h — output formats either side of the hot path: the
// reference only writes P3 text PPM to stdout (rt_in_one_weekend/color.h:14-28,
// accelerated-rt-cuda/final.cu:223-232). Here: P3 (same text format), P6 (binary) and PNG
// (8-bit RGB, zlib "stored" blocks: no compression library needed).
#ifndef RTX_IMAGE_IO_H
#define RTX_IMAGE_IO_H
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <ostream>
#include <string>
#include <vector>

namespace rtx {

inline uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
  return ~crc;
}

inline void png_chunk(std::ostream &out, const char type[4], const std::vector<uint8_t> &data) {
  auto be32 = [&](uint32_t v) { uint8_t b[4] = {uint8_t(v >> 24), uint8_t(v >> 16), uint8_t(v >> 8), uint8_t(v)}; out.write((const char *)b, 4); };
  be32((uint32_t)data.size());
  out.write(type, 4);
  if (!data.empty()) out.write((const char *)data.data(), (std::streamsize)data.size());
  uint32_t crc = crc32_update(0, (const uint8_t *)type, 4);
  if (!data.empty()) crc = crc32_update(crc, data.data(), data.size());
  be32(crc);
}

// rgb: [height][width][3], top row first
inline void write_png(std::ostream &out, const uint8_t *rgb, int width, int height) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  out.write((const char *)sig, 8);
  std::vector<uint8_t> ihdr = {uint8_t(width >> 24), uint8_t(width >> 16), uint8_t(width >> 8), uint8_t(width),
                               uint8_t(height >> 24), uint8_t(height >> 16), uint8_t(height >> 8), uint8_t(height),
                               8, 2, 0, 0, 0}; // 8-bit, colour type 2 (RGB)
  png_chunk(out, "IHDR", ihdr);
  // raw scanlines: filter byte 0 + row
  std::vector<uint8_t> raw;
  raw.reserve((size_t)height * (1 + 3 * (size_t)width));
  for (int j = 0; j < height; j++) {
    raw.push_back(0);
    raw.insert(raw.end(), rgb + (size_t)j * width * 3, rgb + (size_t)(j + 1) * width * 3);
  }
  // zlib stream of stored (uncompressed) deflate blocks
  std::vector<uint8_t> z = {0x78, 0x01};
  size_t pos = 0;
  uint32_t a = 1, b = 0; // adler32
  while (pos < raw.size() || raw.empty()) {
    size_t n = std::min<size_t>(65535, raw.size() - pos);
    const bool last = pos + n >= raw.size();
    z.push_back(last ? 1 : 0);
    z.push_back(uint8_t(n & 0xff)); z.push_back(uint8_t(n >> 8));
    z.push_back(uint8_t(~n & 0xff)); z.push_back(uint8_t((~n >> 8) & 0xff));
    for (size_t i = 0; i < n; i++) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    pos += n;
    if (last) break;
  }
  uint32_t ad = (b << 16) | a;
  z.push_back(uint8_t(ad >> 24)); z.push_back(uint8_t(ad >> 16)); z.push_back(uint8_t(ad >> 8)); z.push_back(uint8_t(ad));
  png_chunk(out, "IDAT", z);
  png_chunk(out, "IEND", {});
}

inline bool write_png_file(const std::string &path, const uint8_t *rgb, int width, int height) {
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  write_png(f, rgb, width, height);
  return (bool)f;
}

} // namespace rtx
#endif
