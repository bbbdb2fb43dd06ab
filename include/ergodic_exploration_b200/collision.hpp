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
 *   DynamicWindow::control(grid, x0, vb, vref | xt_ref)       dynamic_window.cpp:93-187
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
#include <cmath>
#include <stdexcept>
#include <tuple>
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
/** @brief DynamicWindow (dynamic_window.hpp:60-172) evaluated on the GPU, one robot or a batch */
class DynamicWindow
{
public:
  /** @brief same argument order as dynamic_window.hpp:60-65 */
  DynamicWindow(const Collision& collision, double dt, double horizon, double acc_dt, double acc_lim_x,
                double acc_lim_y, double acc_lim_th, double max_vel_x, double min_vel_x, double max_vel_y,
                double min_vel_y, double max_rot_vel, double min_rot_vel, unsigned int vx_samples,
                unsigned int vy_samples, unsigned int vth_samples)
    : collision_(collision)
    , cfg_{ dt,        horizon,   acc_dt,    acc_lim_x,   acc_lim_y,   acc_lim_th, max_vel_x,  min_vel_x,
            max_vel_y, min_vel_y, max_rot_vel, min_rot_vel, vx_samples, vy_samples, vth_samples }
  {
  }

  /** @brief dynamic_window.cpp:93-139: (collision-free twist found, optimal twist) */
  std::tuple<bool, arma::vec> control(const DeviceGrid& grid, const arma::vec& x0, const arma::vec& vb,
                                      const arma::vec& vref) const
  {
    if (x0.n_elem != 3 || vb.n_elem != 3 || vref.n_elem != 3) throw std::logic_error("DynamicWindow::control: 3-vectors");
    int found = 0;
    arma::vec u(3);
    check(eb_dwa_control_twist_host(grid.handle(), &collision_.config(), &cfg_, x0.memptr(), vb.memptr(),
                                    vref.memptr(), 1, &found, u.memptr(), nullptr));
    return std::make_tuple(found != 0, u);
  }

  /** @brief dynamic_window.cpp:141-187: follow a reference trajectory (3 x n, time step dt_ref) */
  std::tuple<bool, arma::vec> control(const DeviceGrid& grid, const arma::vec& x0, const arma::vec& vb,
                                      const arma::mat& xt_ref, double dt_ref) const
  {
    if (x0.n_elem != 3 || vb.n_elem != 3 || xt_ref.n_rows != 3 || xt_ref.n_cols < 1)
      throw std::logic_error("DynamicWindow::control: x0, vb 3-vectors and xt_ref 3 x n");
    int found = 0;
    arma::vec u(3);
    check(eb_dwa_control_traj_host(grid.handle(), &collision_.config(), &cfg_, x0.memptr(), vb.memptr(),
                                   xt_ref.memptr(), static_cast<int>(xt_ref.n_cols), 0, dt_ref, 1, &found, u.memptr(),
                                   nullptr));
    return std::make_tuple(found != 0, u);
  }

  /** @brief batch: x0, vb, vref are 3 x B; returns the B found flags and the 3 x B optimal twists */
  std::tuple<std::vector<int>, arma::mat> controlBatch(const DeviceGrid& grid, const arma::mat& x0, const arma::mat& vb,
                                                      const arma::mat& vref) const
  {
    if (x0.n_rows != 3 || vb.n_rows != 3 || vref.n_rows != 3 || x0.n_cols != vb.n_cols || x0.n_cols != vref.n_cols)
      throw std::logic_error("DynamicWindow::controlBatch: x0, vb, vref must be 3 x B");
    std::vector<int> found(x0.n_cols);
    arma::mat u(3, x0.n_cols);
    check(eb_dwa_control_twist_host(grid.handle(), &collision_.config(), &cfg_, x0.memptr(), vb.memptr(),
                                    vref.memptr(), static_cast<int>(x0.n_cols), found.data(), u.memptr(), nullptr));
    return std::make_tuple(found, u);
  }

  double timeStep() const { return cfg_.dt; }
  double horizon() const { return cfg_.horizon; }
  unsigned int steps() const { return static_cast<unsigned int>(std::abs(cfg_.horizon / cfg_.dt)); }

private:
  Collision collision_;
  eb_dwa cfg_;
};
}  // namespace b200
}  // namespace ergodic_exploration
#endif
