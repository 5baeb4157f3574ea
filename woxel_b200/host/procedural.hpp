// procedural.hpp -- procedural VDB345 scenes for the benchmark configurations of BASELINE.json
// (narrow-band sphere / torus level sets, fractal-noise fog).  Not part of the reference (its
// scenes are .vdb files); definitions follow SURVEY.md section 8(d), configs 3 and 4.
#pragma once
#include <cstdint>

#include "vdb.hpp"

namespace woxel::procedural {

// Voxel (i,j,k) is active iff | |c| - radius | <= band, c = (i+0.5, j+0.5, k+0.5), inside [-half, half)^3.
vdb::VDB345 sphere_shell(int32_t half, double radius, double band);
// Active iff | sqrt((sqrt(cx^2+cz^2) - major)^2 + cy^2) - minor | <= band.
vdb::VDB345 torus_shell(int32_t half, double major, double minor, double band);
// Active iff fbm(c / 256) > tau: value noise, 5 octaves, lacunarity 2, gain 0.5, integer-hash
// lattice seeded 0x9E3779B9, evaluated in f64.  *occupancy (optional) receives the active fraction.
vdb::VDB345 fbm_fog(int32_t half, double tau, double* occupancy);
double fbm(double x, double y, double z);

}  // namespace woxel::procedural
