// frame_dump.cpp -- frame dumps of the capture step: the recorder of the reference pipes RGB frames into an ffmpeg
// child process (src/render/recorder.rs:67-105); here a frame becomes a file directly.  PPM (P6) is the raw dump; PNG
// (8-bit RGB, one zlib stream through the system zlib, filter 0 on every scanline) is the one every viewer opens.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "render.hpp"

namespace woxel::render {

namespace {
struct File {
  FILE* f;
  explicit File(const char* path) : f(fopen(path, "wb")) {
    if (!f) throw std::runtime_error(std::string("cannot open ") + path + " for writing");
  }
  ~File() {
    if (f) fclose(f);
  }
  void write(const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("short write");
  }
  void close() {
    FILE* g = f;
    f = nullptr;
    if (fclose(g) != 0) throw std::runtime_error("close failed");
  }
};

void be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24), p[1] = (uint8_t)(v >> 16), p[2] = (uint8_t)(v >> 8), p[3] = (uint8_t)v; }

void chunk(File& out, const char type[4], const uint8_t* data, size_t n) {
  uint8_t head[8];
  be32(head, (uint32_t)n);
  for (int i = 0; i < 4; ++i) head[4 + i] = (uint8_t)type[i];
  uLong crc = crc32(0L, head + 4, 4);
  if (n) crc = crc32(crc, data, (uInt)n);
  uint8_t tail[4];
  be32(tail, (uint32_t)crc);
  out.write(head, 8);
  out.write(data, n);
  out.write(tail, 4);
}
}  // namespace

void write_ppm(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
  if (!path || !rgb || !width || !height) throw std::invalid_argument("write_ppm: empty frame");
  File out(path);
  char head[64];
  const int n = snprintf(head, sizeof head, "P6\n%u %u\n255\n", width, height);
  out.write(head, (size_t)n);
  out.write(rgb, (size_t)width * height * 3);
  out.close();
}

void write_png(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
  if (!path || !rgb || !width || !height) throw std::invalid_argument("write_png: empty frame");
  const size_t row = (size_t)width * 3;
  if ((uint64_t)(row + 1) * height > 0x7fffffffull) throw std::invalid_argument("write_png: frame too large for one IDAT chunk");
  std::vector<uint8_t> raw((row + 1) * height);
  for (uint32_t y = 0; y < height; ++y) {
    raw[(row + 1) * y] = 0;  // filter type None
    std::copy(rgb + row * y, rgb + row * (y + 1), raw.begin() + (row + 1) * y + 1);
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<uint8_t> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("write_png: zlib failed");
  File out(path);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  out.write(sig, 8);
  uint8_t ihdr[13];
  be32(ihdr, width), be32(ihdr + 4, height);
  ihdr[8] = 8, ihdr[9] = 2, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;  // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
  chunk(out, "IHDR", ihdr, 13);
  // the frame is linear light passed through the recorder's linear_to_srgb: say so (sRGB chunk, perceptual intent)
  const uint8_t intent = 0;
  chunk(out, "sRGB", &intent, 1);
  chunk(out, "IDAT", z.data(), zlen);
  chunk(out, "IEND", nullptr, 0);
  out.close();
}

}  // namespace woxel::render
