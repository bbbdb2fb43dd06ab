// collision_kernels.cuh -- batched occupancy-grid collision checks (sm_100a).
//
// Replaces, for a batch of independent poses / candidate twists sharing one
// occupancy grid, Collision::collisionCheck (collision.cpp:126-143: search
// :150-167, bresenhamCircle :169-215, checkCell :217-244, GridMap::world2Grid
// grid.cpp:143-160, getCell :177-184) and validate_control (numerics.hpp:312-330:
// integrate_twist :273-298 + normalize_angle_PI :77-89 per step) -- the call the
// exploration loop makes right after control() on every tick (exploration.hpp:238).
//
// This is integer / byte gather work: one thread per instance walks the reference's
// Bresenham circles and probes int8 cells through the read-only path (the map is
// shared by the whole batch and stays L2-resident).  Two things differ from a
// transliteration, neither changes a result:
//  * Radius pruning.  checkCell reports a collision only for cells with
//    dx^2 + dy^2 <= r_col^2, and every cell this circle walk emits for radius r
//    lies farther than r - 1/2 from the centre (checked exhaustively for
//    r < kPruneVerified in tests/test_collision_cpu.py), so circles with
//    r > r_col can never hit: the search stops at min(r_max, r_col) instead of
//    r_max.  The closest-obstacle book-keeping of checkCell (cfg.sqrd_obs, dx, dy)
//    feeds minDistance / minDirection only and is not computed.
//  * world2Grid casts floor() to unsigned (undefined for poses left of / below
//    the map); the x86-64 behaviour -- wrap-around, i.e. the signed floor -- is
//    what the CPU checker under tests restates and what is implemented here.
// The pose chain of validate_control uses explicitly rounded multiplies / adds in
// the reference's association order; sin / cos come from the CUDA library, so a
// pose can differ from glibc's by an ulp or two (a cell index could differ only
// for a pose within ~1e-15 of a cell edge).
#pragma once

#include "common.cuh"

namespace eb
{
constexpr int kPruneVerified = 3000;  // radii for which the r - 1/2 bound has been checked exhaustively

struct GridView
{
  const signed char* data;  // [ysize][xsize]
  unsigned int xsize, ysize;
  double resolution, xmin, ymin;
};

struct CollisionParams
{
  GridView g;
  int r_bnd, r_col, r_max;  // collision.cpp:130-133
  double occupied_threshold;
  int B;
  // validate_control
  int steps;
  double dt;
  const double* x0;  // [B][3] start poses (validate) or the poses themselves (check)
  const double* u;   // [B][3] twists (validate only)
  int* out;          // [B]: collision_check 1 = collision; validate_control 1 = collision free
};

// numerics.hpp:77-89 with every operation explicitly rounded (no FMA contraction)
__device__ __forceinline__ double normalize_angle_pi_rn(double rad)
{
  const double shifted = __dadd_rn(rad, kPi);
  const double q = floor(__ddiv_rn(shifted, __dmul_rn(2.0, kPi)));
  rad = __dsub_rn(shifted, __dmul_rn(__dmul_rn(q, 2.0), kPi));
  if (rad < 0.0) rad = __dadd_rn(rad, __dmul_rn(2.0, kPi));
  return __dsub_rn(rad, kPi);
}

// Collision::checkCell (:217-244) without the closest-obstacle book-keeping
__device__ __forceinline__ bool check_cell(const GridView& g, double thr, int cx, int cy, int r_col, unsigned int cj,
                                           unsigned int ci)
{
  if (!(ci <= g.ysize - 1u && cj <= g.xsize - 1u)) return false;  // GridMap::gridBounds grid.cpp:96-100
  const double cell = (double)__ldg(g.data + (size_t)ci * g.xsize + cj) / 100.0;  // getCell grid.cpp:177-184
  if (cell < thr) return false;
  const unsigned int dj = (unsigned int)cx - cj, di = (unsigned int)cy - ci;
  const int sqrd_obs = (int)(dj * dj + di * di);
  return sqrd_obs <= r_col * r_col;
}

// Collision::collisionCheck (:126-143) for one pose
__device__ __forceinline__ bool collision_check_pose(const CollisionParams& p, double px, double py)
{
  const GridView& g = p.g;
  // GridMap::world2Grid grid.cpp:143-160
  unsigned int j = (unsigned int)(long long)floor((px - g.xmin) / g.resolution);
  unsigned int i = (unsigned int)(long long)floor((py - g.ymin) / g.resolution);
  if (j == g.xsize) j--;
  if (i == g.ysize) i--;
  const int cx = (int)j, cy = (int)i;
  // Collision::search (:150-167), pruned to the radii that can satisfy sqrd_obs <= r_col^2
  const int r_last = (p.r_col + 1 < kPruneVerified) ? min(p.r_max, p.r_col) : p.r_max;
  for (int r0 = p.r_bnd; r0 <= r_last; r0++)
  {
    // Collision::bresenhamCircle (:169-215)
    int x = -r0, y = 0, err = 2 - 2 * r0;
    while (x < 0)
    {
      if (check_cell(g, p.occupied_threshold, cx, cy, p.r_col, (unsigned int)(cx - x), (unsigned int)(cy + y)) ||
          check_cell(g, p.occupied_threshold, cx, cy, p.r_col, (unsigned int)(cx - y), (unsigned int)(cy - x)) ||
          check_cell(g, p.occupied_threshold, cx, cy, p.r_col, (unsigned int)(cx + x), (unsigned int)(cy - y)) ||
          check_cell(g, p.occupied_threshold, cx, cy, p.r_col, (unsigned int)(cx + y), (unsigned int)(cy + x)))
        return true;
      const int r = err;
      if (r <= y)
      {
        y++;
        err += 2 * y + 1;
      }
      if (r > x || err > y)
      {
        x++;
        err += 2 * x + 1;
      }
    }
  }
  return false;
}

__global__ void __launch_bounds__(128) collision_check_kernel(const CollisionParams p)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B) return;
  p.out[i] = collision_check_pose(p, p.x0[(size_t)i * 3 + 0], p.x0[(size_t)i * 3 + 1]) ? 1 : 0;
}

// validate_control (numerics.hpp:312-330): constant twist integrated for `steps` steps,
// the pose checked after every step; 1 = collision free
__global__ void __launch_bounds__(128) validate_control_kernel(const CollisionParams p)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B) return;
  double x = p.x0[(size_t)i * 3 + 0], y = p.x0[(size_t)i * 3 + 1], th = p.x0[(size_t)i * 3 + 2];
  const double u0 = p.u[(size_t)i * 3 + 0], u1 = p.u[(size_t)i * 3 + 1], u2 = p.u[(size_t)i * 3 + 2];
  // integrate_twist (:273-298): the body-frame displacement of one step is the same every step
  double d0, d1, d2;
  if (fabs(u2 - 0.0) < 1.0e-12)
  {
    d0 = __dmul_rn(u0, p.dt);
    d1 = __dmul_rn(u1, p.dt);
    d2 = 0.0;
  }
  else
  {
    const double vb0 = __dmul_rn(u0, p.dt), vb1 = __dmul_rn(u1, p.dt), vb2 = __dmul_rn(u2, p.dt);
    double s, c;
    sincos(vb2, &s, &c);
    d0 = __ddiv_rn(__dadd_rn(__dmul_rn(vb0, s), __dmul_rn(vb1, __dsub_rn(c, 1.0))), vb2);
    d1 = __ddiv_rn(__dadd_rn(__dmul_rn(vb1, s), __dmul_rn(vb0, __dsub_rn(1.0, c))), vb2);
    d2 = vb2;
  }
  int ok = 1;
  for (int k = 0; k < p.steps; k++)
  {
    double s, c;
    sincos(th, &s, &c);
    // x + transform2d(theta) * dqb, rows summed in column order (the third column is zero)
    x = __dadd_rn(x, __dadd_rn(__dmul_rn(c, d0), __dmul_rn(-s, d1)));
    y = __dadd_rn(y, __dadd_rn(__dmul_rn(s, d0), __dmul_rn(c, d1)));
    th = normalize_angle_pi_rn(__dadd_rn(th, d2));
    if (collision_check_pose(p, x, y))
    {
      ok = 0;
      break;
    }
  }
  p.out[i] = ok;
}
}  // namespace eb
