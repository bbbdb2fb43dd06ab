/**
 * @file collision.hpp
 * @brief Header-only C++ adapter for the batched occupancy-grid collision checks
 *        of libergodic_b200 (include/ergodic_b200.h, eb_grid_* / eb_collision_check_* /
 *        eb_validate_control_*).
 *
 * Mirrors, in namespace ergodic_exploration::b200, what the exploration loop calls right
 * after ErgodicControl::control() on every tick (exploration.hpp:238):
 *   Collision::collisionCheck(grid, pose)                     collision.cpp:126-143
 *   validate_control(collision, grid, x0, u, dt, horizon)     numerics.hpp:312-330
 * for one pose / twist (the reference's signatures) or for a 3 x B batch sharing one map.
 * The grid type is anything with the reference GridMap's getters
 * (gridData(), xsize(), ysize(), resolution(), xmin(), ymin(); grid.hpp:201-255).
 * The reference's Collision keeps its radii private (collision.hpp:159-163), so the
 * four constructor arguments are given to this class directly; it throws where the
 * reference constructor throws (collision.cpp:52-63).  Armadillo in, Armadillo / std out.
 */
#ifndef ERGODIC_EXPLORATION_B200_COLLISION_HPP
#define ERGODIC_EXPLORATION_B200_COLLISION_HPP

#include <memory>
#include <stdexcept>
#include <vector>

#include <armadillo>

#include <ergodic_b200.h>
#include <ergodic_exploration_b200/ergodic_control.hpp>

namespace ergodic_exploration
{
namespace b200
{
/** @brief An occupancy grid resident on the GPU (copy of a GridMap's cells and geometry) */
class DeviceGrid
{
public:
  template <class GridT>
  explicit DeviceGrid(const GridT& grid)
    : xsize_(grid.xsize()), ysize_(grid.ysize()), resolution_(grid.resolution()), xmin_(grid.xmin()), ymin_(grid.ymin())
  {
    eb_grid* g = nullptr;
    check(eb_grid_create(default_device(), reinterpret_cast<const signed char*>(grid.gridData().data()), xsize_,
                         ysize_, resolution_, xmin_, ymin_, &g));
    h_.reset(g, eb_grid_destroy);
  }

  /** @brief GridMap::update (grid.cpp:84-94): new cells; a new geometry re-creates the device copy */
  template <class GridT>
  void update(const GridT& grid)
  {
    if (grid.xsize() != xsize_ || grid.ysize() != ysize_ || grid.resolution() != resolution_ ||
        grid.xmin() != xmin_ || grid.ymin() != ymin_)
    {
      *this = DeviceGrid(grid);
      return;
    }
    check(eb_grid_update(h_.get(), reinterpret_cast<const signed char*>(grid.gridData().data())));
  }

  eb_grid* handle() const { return h_.get(); }

private:
  unsigned int xsize_, ysize_;
  double resolution_, xmin_, ymin_;
  std::shared_ptr<eb_grid> h_;
};

/** @brief Collision (collision.hpp:85-165) evaluated on the GPU, one pose or a batch */
class Collision
{
public:
  Collision(double boundary_radius, double search_radius, double obstacle_threshold, double occupied_threshold)
    : cfg_{ boundary_radius, search_radius, obstacle_threshold, occupied_threshold }
  {
    if (search_radius < boundary_radius)  // collision.cpp:52-56
      throw std::invalid_argument("Search radius must be at least the same size as the boundary radius");
    if (occupied_threshold > 100.0 || occupied_threshold < 0.0)  // collision.cpp:58-62
      throw std::invalid_argument("Occupied threshold must be between 0 and 100");
  }

  double totalPadding() const { return cfg_.boundary_radius + cfg_.obstacle_threshold; }  // collision.cpp:145-148

  /** @brief collision.cpp:126-143 */
  bool collisionCheck(const DeviceGrid& grid, const arma::vec& pose) const
  {
    if (pose.n_elem != 3) throw std::logic_error("collisionCheck: pose must have 3 elements");
    int hit = 0;
    check(eb_collision_check_host(grid.handle(), &cfg_, pose.memptr(), 1, &hit));
    return hit != 0;
  }

  /** @brief batch: poses is 3 x B; returns B flags, 1 = collision */
  std::vector<int> collisionCheck(const DeviceGrid& grid, const arma::mat& poses) const
  {
    if (poses.n_rows != 3) throw std::logic_error("collisionCheck: poses must be 3 x B");
    std::vector<int> hit(poses.n_cols);
    check(eb_collision_check_host(grid.handle(), &cfg_, poses.memptr(), static_cast<int>(poses.n_cols), hit.data()));
    return hit;
  }

  const eb_collision& config() const { return cfg_; }

private:
  eb_collision cfg_;
};

/** @brief numerics.hpp:312-330: true if the constant twist is collision free over the horizon */
inline bool validate_control(const Collision& collision, const DeviceGrid& grid, const arma::vec& x0,
                             const arma::vec& u, double dt, double horizon)
{
  if (x0.n_elem != 3 || u.n_elem != 3) throw std::logic_error("validate_control: x0 and u must have 3 elements");
  int valid = 0;
  check(eb_validate_control_host(grid.handle(), &collision.config(), x0.memptr(), u.memptr(), 1, dt, horizon, &valid));
  return valid != 0;
}

/** @brief batch: x0 and u are 3 x B (B start poses / candidate twists on one map) */
inline std::vector<int> validate_control(const Collision& collision, const DeviceGrid& grid, const arma::mat& x0,
                                         const arma::mat& u, double dt, double horizon)
{
  if (x0.n_rows != 3 || u.n_rows != 3 || x0.n_cols != u.n_cols)
    throw std::logic_error("validate_control: x0 and u must be 3 x B");
  std::vector<int> valid(x0.n_cols);
  check(eb_validate_control_host(grid.handle(), &collision.config(), x0.memptr(), u.memptr(),
                                 static_cast<int>(x0.n_cols), dt, horizon, valid.data()));
  return valid;
}
}  // namespace b200
}  // namespace ergodic_exploration
#endif
