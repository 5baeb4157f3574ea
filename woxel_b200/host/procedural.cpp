// procedural.cpp -- see procedural.hpp.
#include "procedural.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <thread>

namespace woxel::procedural {

namespace {

// distance range of |x| over the interval [lo, hi]
inline void abs_range(double lo, double hi, double& mn, double& mx) {
  mx = std::max(std::fabs(lo), std::fabs(hi));
  mn = (lo <= 0.0 && hi >= 0.0) ? 0.0 : std::min(std::fabs(lo), std::fabs(hi));
}

// Walks the 8^3 leaf cells of [-half, half)^3; `maybe(o, edge)` conservatively tells whether the cube
// of voxel centres with corner voxel o and `edge` voxels per side can hold an active voxel;
// `active(x,y,z)` is the exact per-voxel predicate.
template <class Maybe, class Active>
vdb::VDB345 build(int32_t half, Maybe&& maybe, Active&& active) {
  vdb::VDB345 out;
  struct Found {
    vdb::GlobalCoordinates origin;
    uint64_t mask[8];
  };
  // coarse blocks of 128^3 (one N4) are distributed over threads; results are merged in block order
  std::vector<vdb::GlobalCoordinates> blocks;
  for (int32_t x = -half; x < half; x += 128)
    for (int32_t y = -half; y < half; y += 128)
      for (int32_t z = -half; z < half; z += 128)
        if (maybe(vdb::GlobalCoordinates{x, y, z}, 128)) blocks.push_back({x, y, z});
  std::vector<std::vector<Found>> per_block(blocks.size());
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const size_t b = next.fetch_add(1);
      if (b >= blocks.size()) break;
      const auto o = blocks[b];
      for (int32_t lx = 0; lx < 128; lx += 8)
        for (int32_t ly = 0; ly < 128; ly += 8)
          for (int32_t lz = 0; lz < 128; lz += 8) {
            const vdb::GlobalCoordinates lo = {o[0] + lx, o[1] + ly, o[2] + lz};
            if (lo[0] >= half || lo[1] >= half || lo[2] >= half) continue;
            if (!maybe(lo, 8)) continue;
            Found f;
            f.origin = lo;
            bool any = false;
            for (auto& w : f.mask) w = 0;
            for (uint32_t off = 0; off < 512; ++off) {
              const int32_t x = lo[0] + (int32_t)(off >> 6), y = lo[1] + (int32_t)((off >> 3) & 7), z = lo[2] + (int32_t)(off & 7);
              if (x < half && y < half && z < half && active(x, y, z)) f.mask[off >> 6] |= 1ull << (off & 63), any = true;
            }
            if (any) per_block[b].push_back(f);
          }
    }
  };
  const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  for (const auto& v : per_block)
    for (const Found& f : v) out.add_leaf(f.origin, f.mask);
  return out;
}

}  // namespace

vdb::VDB345 sphere_shell(int32_t half, double radius, double band) {
  auto maybe = [=](vdb::GlobalCoordinates o, int32_t edge) {
    double mn2 = 0, mx2 = 0;
    for (int a = 0; a < 3; ++a) {
      double mn, mx;
      abs_range(o[a] + 0.5, o[a] + edge - 0.5, mn, mx);
      mn2 += mn * mn, mx2 += mx * mx;
    }
    return std::sqrt(mx2) >= radius - band && std::sqrt(mn2) <= radius + band;
  };
  auto active = [=](int32_t x, int32_t y, int32_t z) {
    const double cx = x + 0.5, cy = y + 0.5, cz = z + 0.5;
    return std::fabs(std::sqrt(cx * cx + cy * cy + cz * cz) - radius) <= band;
  };
  return build(half, maybe, active);
}

vdb::VDB345 torus_shell(int32_t half, double major, double minor, double band) {
  auto maybe = [=](vdb::GlobalCoordinates o, int32_t edge) {
    double xmn, xmx, ymn, ymx, zmn, zmx;
    abs_range(o[0] + 0.5, o[0] + edge - 0.5, xmn, xmx);
    abs_range(o[1] + 0.5, o[1] + edge - 0.5, ymn, ymx);
    abs_range(o[2] + 0.5, o[2] + edge - 0.5, zmn, zmx);
    // ring distance q = sqrt(cx^2+cz^2) - major ranges over [qlo, qhi]
    const double qlo = std::sqrt(xmn * xmn + zmn * zmn) - major, qhi = std::sqrt(xmx * xmx + zmx * zmx) - major;
    double qmn, qmx;
    abs_range(qlo, qhi, qmn, qmx);
    const double dmn = std::sqrt(qmn * qmn + ymn * ymn), dmx = std::sqrt(qmx * qmx + ymx * ymx);
    return dmx >= minor - band && dmn <= minor + band;
  };
  auto active = [=](int32_t x, int32_t y, int32_t z) {
    const double cx = x + 0.5, cy = y + 0.5, cz = z + 0.5;
    const double q = std::sqrt(cx * cx + cz * cz) - major;
    return std::fabs(std::sqrt(q * q + cy * cy) - minor) <= band;
  };
  return build(half, maybe, active);
}

namespace {
inline uint32_t hash3(int32_t x, int32_t y, int32_t z) {
  uint32_t h = 0x9E3779B9u;
  h ^= (uint32_t)x * 0x85EBCA6Bu, h = (h << 13) | (h >> 19), h *= 0xC2B2AE35u;
  h ^= (uint32_t)y * 0x27D4EB2Fu, h = (h << 13) | (h >> 19), h *= 0xC2B2AE35u;
  h ^= (uint32_t)z * 0x165667B1u, h = (h << 13) | (h >> 19), h *= 0xC2B2AE35u;
  h ^= h >> 16, h *= 0x85EBCA6Bu, h ^= h >> 13, h *= 0xC2B2AE35u, h ^= h >> 16;
  return h;
}
inline double lattice(int32_t x, int32_t y, int32_t z) { return (double)hash3(x, y, z) * (1.0 / 4294967296.0); }
inline double smooth(double t) { return t * t * (3.0 - 2.0 * t); }
double value_noise(double x, double y, double z) {
  const double fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
  const int32_t ix = (int32_t)fx, iy = (int32_t)fy, iz = (int32_t)fz;
  const double tx = smooth(x - fx), ty = smooth(y - fy), tz = smooth(z - fz);
  double c[2][2][2];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int d = 0; d < 2; ++d) c[a][b][d] = lattice(ix + a, iy + b, iz + d);
  auto lerp = [](double a, double b, double t) { return a + (b - a) * t; };
  const double x00 = lerp(c[0][0][0], c[1][0][0], tx), x10 = lerp(c[0][1][0], c[1][1][0], tx);
  const double x01 = lerp(c[0][0][1], c[1][0][1], tx), x11 = lerp(c[0][1][1], c[1][1][1], tx);
  return lerp(lerp(x00, x10, ty), lerp(x01, x11, ty), tz);
}
}  // namespace

double fbm(double x, double y, double z) {
  double sum = 0.0, amp = 0.5, norm = 0.0;
  for (int o = 0; o < 5; ++o) {
    sum += amp * value_noise(x, y, z);
    norm += amp;
    x *= 2.0, y *= 2.0, z *= 2.0;
    amp *= 0.5;
  }
  return sum / norm;
}

vdb::VDB345 fbm_fog(int32_t half, double tau, double* occupancy) {
  auto maybe = [](vdb::GlobalCoordinates, int32_t) { return true; };
  auto active = [=](int32_t x, int32_t y, int32_t z) { return fbm((x + 0.5) / 256.0, (y + 0.5) / 256.0, (z + 0.5) / 256.0) > tau; };
  vdb::VDB345 v = build(half, maybe, active);
  if (occupancy) *occupancy = (double)v.count_leaf_values() / (8.0 * (double)half * (double)half * (double)half);
  return v;
}

}  // namespace woxel::procedural
