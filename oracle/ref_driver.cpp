// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI driver around the UNMODIFIED reference sources, which are compiled
// where they lie under /root/reference (never copied) against the test-only
// Armadillo/ROS header shim in oracle/shim.  The result, oracle/_ref/
// libergodic_ref.so, is the "compiled reference": it pins the plain-C
// restatement (oracle/ergodic_oracle.c), generates the golden vectors under
// tests/golden/, and is the preferred CPU baseline in bench.py.
//
// Reference private members (ut_, phik_, basis_, ...) are reached with
// `#define private public` placed AFTER every standard header has been
// included, so only the reference's own classes are affected.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include <armadillo>
#include <ros/stub_msgs.h>

#define private public
#include <ergodic_exploration/ergodic_control.hpp>
#include <ergodic_exploration/models/cart.hpp>
#include <ergodic_exploration/models/omni.hpp>
#include <ergodic_exploration/dynamic_window.hpp>
#undef private

using arma::mat;
using arma::vec;
namespace ee = ergodic_exploration;

namespace
{
// configTarget prints progress to stdout (ergodic_control.hpp:379,415) and the
// replay buffer warns when full (buffer.cpp:61).  The driver discards C++
// std::cout for the life of the library (stateless sink: safe with the
// multi-threaded CPU baseline).
struct NullBuf : std::streambuf
{
  int overflow(int c) override { return c; }
  std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
NullBuf g_null;
struct QuietInstaller
{
  QuietInstaller() { std::cout.rdbuf(&g_null); }
} g_quiet_installer;
struct QuietCout
{
};

vec v3(const double* p) { return vec({ p[0], p[1], p[2] }); }
void out_mat(const mat& m, double* dst) { std::memcpy(dst, m.memptr(), sizeof(double) * m.n_elem); }

ee::GridMap make_grid(double xmin, double xmax, double ymin, double ymax, double map_res)
{
  const auto xs = ee::axis_length(xmin, xmax, map_res);
  const auto ys = ee::axis_length(ymin, ymax, map_res);
  return ee::GridMap(xmin, xmax, ymin, ymax, map_res, ee::GridData(size_t(xs) * ys, 0));
}

struct RefController
{
  int model;
  std::unique_ptr<ee::ErgodicControl<ee::models::SimpleCart>> cart;
  std::unique_ptr<ee::ErgodicControl<ee::models::Omni>> omni;
  // by-products of the last ref_control_trace()
  vec ck;
  mat edx, bdx, rhot, xtf;
};

template <class F>
auto with(RefController* c, F f)
{
  return c->model == 0 ? f(*c->cart) : f(*c->omni);
}

std::string g_err;
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ---- leaf kinematics (test/test_cart.cpp, test/test_omni.cpp) -------------
int ref_model_f(int model, const double* x, const double* u, double* xdot)
{
  try
  {
    const vec r = model == 0 ? ee::models::SimpleCart()(v3(x), v3(u)) : ee::models::Omni()(v3(x), v3(u));
    out_mat(r, xdot);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}
void ref_model_fdx(int model, const double* x, const double* u, double* A)
{
  out_mat(model == 0 ? ee::models::SimpleCart().fdx(v3(x), v3(u)) : ee::models::Omni().fdx(v3(x), v3(u)), A);
}
void ref_model_fdu(int model, const double* x, double* B)
{
  out_mat(model == 0 ? ee::models::SimpleCart().fdu(v3(x)) : ee::models::Omni().fdu(v3(x)), B);
}
void ref_cart(double r, double b, const double* x, const double* u, double* xdot, double* A, double* B,
              double* twist)
{
  const ee::models::Cart m(r, b);
  const vec uu({ u[0], u[1] });
  out_mat(m(v3(x), uu), xdot);
  out_mat(m.fdx(v3(x), uu), A);
  out_mat(m.fdu(v3(x)), B);
  out_mat(m.wheels2Twist(uu), twist);
}
void ref_mecanum(double r, double bx, double by, const double* x, const double* u, double* xdot, double* A,
                 double* B, double* twist)
{
  const ee::models::Mecanum m(r, bx, by);
  const vec uu({ u[0], u[1], u[2], u[3] });
  out_mat(m(v3(x), uu), xdot);
  out_mat(m.fdx(v3(x), uu), A);
  out_mat(m.fdu(v3(x)), B);
  out_mat(m.wheels2Twist(uu), twist);
}

// ---- numerics / integrator ------------------------------------------------
double ref_normalize_angle_pi(double r) { return ee::normalize_angle_PI(r); }
void ref_integrate_twist(const double* x, const double* u, double dt, double* out)
{
  out_mat(ee::integrate_twist(v3(x), v3(u), dt), out);
}
int ref_rk4_forward(int model, double dt, double horizon, const double* x0, const double* ut, double* xt)
{
  try
  {
    const unsigned steps = static_cast<unsigned>(std::abs(horizon / dt));
    mat u(3, steps);
    std::memcpy(u.memptr(), ut, sizeof(double) * 3 * steps);
    const ee::RungeKutta rk(dt);
    const mat r = model == 0 ? rk.solve(ee::models::SimpleCart(), v3(x0), u, horizon) :
                               rk.solve(ee::models::Omni(), v3(x0), u, horizon);
    out_mat(r, xt);
    return int(steps);
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}
int ref_rk4_forward_cart(double r, double b, double dt, double horizon, const double* x0, const double* ut,
                         double* xt)
{
  const unsigned steps = static_cast<unsigned>(std::abs(horizon / dt));
  mat u(2, steps);
  std::memcpy(u.memptr(), ut, sizeof(double) * 2 * steps);
  const ee::RungeKutta rk(dt);
  out_mat(rk.solve(ee::models::Cart(r, b), v3(x0), u, horizon), xt);
  return int(steps);
}

int ref_rk4_forward_mecanum(double r, double bx, double by, double dt, double horizon, const double* x0,
                            const double* ut, double* xt)
{
  const unsigned steps = static_cast<unsigned>(std::abs(horizon / dt));
  mat u(4, steps);
  std::memcpy(u.memptr(), ut, sizeof(double) * 4 * steps);
  const ee::RungeKutta rk(dt);
  out_mat(rk.solve(ee::models::Mecanum(r, bx, by), v3(x0), u, horizon), xt);
  return int(steps);
}
double ref_entropy(double p) { return ee::entropy(p); }
// entropy(getCell(idx)) for every cell of a GridMap built from the raw int8 data (grid.cpp:177-184, numerics.hpp:164-179)
void ref_entropy_grid(const signed char* data, unsigned xsize, unsigned ysize, double res, double* out)
{
  const std::vector<int8_t> cells(data, data + (size_t)xsize * ysize);
  const ee::GridMap grid(0.0, xsize * res, 0.0, ysize * res, res, cells);
  for (unsigned i = 0; i < xsize * ysize; i++) out[i] = ee::entropy(grid.getCell(i));
}

// ---- basis / target -------------------------------------------------------
void ref_basis_tables(int nb, long long* k, double* lamdak)
{
  const ee::Basis b(1.0, 1.0, nb);
  std::memcpy(k, b.k_.memptr(), sizeof(long long) * b.k_.n_elem);
  out_mat(b.lamdak_, lamdak);
}
void ref_fourier_basis(double lx, double ly, int nb, const double* x, double* fk)
{
  out_mat(ee::Basis(lx, ly, nb).fourierBasis(vec({ x[0], x[1] })), fk);
}
void ref_grad_fourier_basis(double lx, double ly, int nb, const double* x, double* dfk)
{
  out_mat(ee::Basis(lx, ly, nb).gradFourierBasis(vec({ x[0], x[1] })), dfk);
}
void ref_traj_coeff(double lx, double ly, int nb, const double* xt, int ld, int ncols, double* ck)
{
  mat m(ld, ncols);
  std::memcpy(m.memptr(), xt, sizeof(double) * size_t(ld) * ncols);
  out_mat(ee::Basis(lx, ly, nb).trajCoeff(m), ck);
}
void ref_spatial_coeff(double lx, double ly, int nb, const double* phi_vals, const double* phi_grid,
                       long long G, double* phik)
{
  vec pv(G);
  mat pg(2, G);
  std::memcpy(pv.memptr(), phi_vals, sizeof(double) * G);
  std::memcpy(pg.memptr(), phi_grid, sizeof(double) * 2 * G);
  out_mat(ee::Basis(lx, ly, nb).spatialCoeff(pv, pg), phik);
}
static ee::Target make_target(int ng, const double* mu, const double* sigma)
{
  ee::GaussianList gl;
  for (int g = 0; g < ng; g++)
    gl.emplace_back(vec({ mu[2 * g], mu[2 * g + 1] }), vec({ sigma[2 * g], sigma[2 * g + 1] }));
  return ee::Target(gl);
}
void ref_target_fill(int ng, const double* mu, const double* sigma, const double* trans,
                     const double* phi_grid, long long G, double* phi_vals)
{
  mat pg(2, G);
  std::memcpy(pg.memptr(), phi_grid, sizeof(double) * 2 * G);
  out_mat(make_target(ng, mu, sigma).fill(vec({ trans[0], trans[1] }), pg), phi_vals);
}

// ---- ErgodicControl -------------------------------------------------------
void* ref_create(int model, double dt, double horizon, double resolution, double expl_weight, int num_basis,
                 long long buffer_size, int batch_size, const double* Rinv, const double* umin,
                 const double* umax)
{
  try
  {
    mat R(3, 3);
    std::memcpy(R.memptr(), Rinv, sizeof(double) * 9);
    const ee::Collision collision(0.7, 1.0, 0.2, 0.8);  // stored, never used by the controller
    auto c = std::make_unique<RefController>();
    c->model = model;
    if (model == 0)
      c->cart = std::make_unique<ee::ErgodicControl<ee::models::SimpleCart>>(
          ee::models::SimpleCart(), collision, dt, horizon, resolution, expl_weight, unsigned(num_basis),
          unsigned(buffer_size), unsigned(batch_size), R, v3(umin), v3(umax));
    else
      c->omni = std::make_unique<ee::ErgodicControl<ee::models::Omni>>(
          ee::models::Omni(), collision, dt, horizon, resolution, expl_weight, unsigned(num_basis),
          unsigned(buffer_size), unsigned(batch_size), R, v3(umin), v3(umax));
    return c.release();
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return nullptr;
  }
}
void ref_destroy(void* h) { delete static_cast<RefController*>(h); }

void ref_set_target(void* h, int ng, const double* mu, const double* sigma)
{
  const ee::Target t = make_target(ng, mu, sigma);
  with(static_cast<RefController*>(h), [&](auto& ec) {
    ec.setTarget(t);
    return 0;
  });
}
void ref_add_state_memory(void* h, const double* x)
{
  QuietCout q;
  with(static_cast<RefController*>(h), [&](auto& ec) {
    ec.addStateMemory(v3(x));
    return 0;
  });
}
long long ref_memory_size(void* h)
{
  return with(static_cast<RefController*>(h), [&](auto& ec) { return (long long)ec.buffer_.memory_.size(); });
}
int ref_steps(void* h)
{
  return with(static_cast<RefController*>(h), [&](auto& ec) { return int(ec.steps_); });
}

// the reference's own control() (ergodic_control.hpp:225-311), untouched
int ref_control(void* h, double xmin, double xmax, double ymin, double ymax, double map_res, const double* x,
                double* u0)
{
  QuietCout q;
  try
  {
    const ee::GridMap grid = make_grid(xmin, xmax, ymin, ymax, map_res);
    const vec u = with(static_cast<RefController*>(h), [&](auto& ec) { return ec.control(grid, v3(x)); });
    out_mat(u, u0);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

// Same sequence of reference member calls as control() (:227-310), issued one
// by one so the intermediates (c_k, edx, bdx, rhot) can be captured.
int ref_control_trace(void* h, double xmin, double xmax, double ymin, double ymax, double map_res,
                      const double* x, double* u0)
{
  QuietCout q;
  RefController* c = static_cast<RefController*>(h);
  try
  {
    const ee::GridMap grid = make_grid(xmin, xmax, ymin, ymax, map_res);
    const vec u = with(c, [&](auto& ec) {
      ec.pose_ = v3(x);
      ec.configTarget(grid);
      ec.ut_.cols(0, ec.ut_.n_cols - 2) = ec.ut_.cols(1, ec.ut_.n_cols - 1);
      ec.ut_.col(ec.ut_.n_cols - 1).fill(0.0);
      const mat traj = ec.rk4_.solve(ec.model_, ec.pose_, ec.ut_, ec.horizon_);
      mat xt_total = ec.buffer_.sampleMemory(traj);
      xt_total.row(0) -= ec.map_pos_(0);
      xt_total.row(1) -= ec.map_pos_(1);
      const mat xt = xt_total.cols(xt_total.n_cols - ec.steps_, xt_total.n_cols - 1);
      c->ck = ec.basis_.trajCoeff(xt_total);
      c->edx = ec.gradErgodicMetric(c->ck, xt);
      c->bdx = ec.gradBarrier(xt);
      c->rhot = ec.rk4_.solve(ec.rhodot_, ec.model_, ec.rhoT_, xt, ec.ut_, c->edx, c->bdx, ec.horizon_);
      c->xtf = xt;
      ec.updateControl(xt, c->rhot);
      return vec(ec.ut_.col(0));
    });
    out_mat(u, u0);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}
void ref_get_last(void* h, double* ck, double* edx, double* bdx, double* rhot, double* xtf)
{
  RefController* c = static_cast<RefController*>(h);
  if (ck) out_mat(c->ck, ck);
  if (edx) out_mat(c->edx, edx);
  if (bdx) out_mat(c->bdx, bdx);
  if (rhot) out_mat(c->rhot, rhot);
  if (xtf) out_mat(c->xtf, xtf);
}
int ref_config_target(void* h, double xmin, double xmax, double ymin, double ymax, double map_res)
{
  QuietCout q;
  const ee::GridMap grid = make_grid(xmin, xmax, ymin, ymax, map_res);
  with(static_cast<RefController*>(h), [&](auto& ec) {
    ec.configTarget(grid);
    return 0;
  });
  return 0;
}
int ref_opt_traj(void* h, double* xt)
{
  try
  {
    const mat t = with(static_cast<RefController*>(h), [&](auto& ec) { return ec.optTraj(); });
    out_mat(t, xt);
    return int(t.n_cols);
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}
void ref_get_ut(void* h, double* ut)
{
  with(static_cast<RefController*>(h), [&](auto& ec) {
    out_mat(ec.ut_, ut);
    return 0;
  });
}
void ref_set_ut(void* h, const double* ut)
{
  with(static_cast<RefController*>(h), [&](auto& ec) {
    std::memcpy(ec.ut_.memptr(), ut, sizeof(double) * ec.ut_.n_elem);
    return 0;
  });
}
void ref_get_phik(void* h, double* phik, double* lxy)
{
  with(static_cast<RefController*>(h), [&](auto& ec) {
    out_mat(ec.phik_, phik);
    if (lxy)
    {
      lxy[0] = ec.basis_.lx_;
      lxy[1] = ec.basis_.ly_;
    }
    return 0;
  });
}

// Batched driver for the CPU baseline: `count` independent controllers, one
// control() each, on the calling thread.
int ref_control_many(void** hs, int count, double xmin, double xmax, double ymin, double ymax, double map_res,
                     const double* x, double* u0)
{
  QuietCout q;
  try
  {
    const ee::GridMap grid = make_grid(xmin, xmax, ymin, ymax, map_res);
    for (int i = 0; i < count; i++)
    {
      const vec u =
          with(static_cast<RefController*>(hs[i]), [&](auto& ec) { return ec.control(grid, v3(x + 3 * i)); });
      out_mat(u, u0 + 3 * i);
    }
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

// ---- occupancy-grid collision checking (collision.cpp, numerics.hpp:312-330) ----
// The grid is rebuilt through the reference's own GridMap constructor
// (grid.cpp:46-61): xmax / ymax are chosen so that axis_length() returns xsize / ysize.
static ee::GridMap occupancy_grid(const signed char* data, unsigned xsize, unsigned ysize, double res, double xmin,
                                  double ymin)
{
  const double xmax = ee::axis_upper(xmin, res, xsize), ymax = ee::axis_upper(ymin, res, ysize);
  return ee::GridMap(xmin, xmax, ymin, ymax, res, ee::GridData(data, data + size_t(xsize) * ysize));
}

int ref_collision_check_many(const signed char* data, unsigned xsize, unsigned ysize, double res, double xmin,
                             double ymin, double boundary_radius, double search_radius, double obstacle_threshold,
                             double occupied_threshold, const double* poses, int count, int* hit)
{
  try
  {
    const ee::GridMap grid = occupancy_grid(data, xsize, ysize, res, xmin, ymin);
    const ee::Collision col(boundary_radius, search_radius, obstacle_threshold, occupied_threshold);
    for (int i = 0; i < count; i++) hit[i] = col.collisionCheck(grid, v3(poses + 3 * i)) ? 1 : 0;
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

int ref_validate_control_many(const signed char* data, unsigned xsize, unsigned ysize, double res, double xmin,
                              double ymin, double boundary_radius, double search_radius, double obstacle_threshold,
                              double occupied_threshold, const double* x0, const double* u, int count, double dt,
                              double horizon, int* valid)
{
  try
  {
    const ee::GridMap grid = occupancy_grid(data, xsize, ysize, res, xmin, ymin);
    const ee::Collision col(boundary_radius, search_radius, obstacle_threshold, occupied_threshold);
    for (int i = 0; i < count; i++)
      valid[i] = ee::validate_control(col, grid, v3(x0 + 3 * i), v3(u + 3 * i), dt, horizon) ? 1 : 0;
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

// ---- DynamicWindow (dynamic_window.cpp): cfg = {dt, horizon, acc_dt, acc_lim_x, acc_lim_y, acc_lim_th,
//      max_vel_x, min_vel_x, max_vel_y, min_vel_y, max_rot_vel, min_rot_vel}, samples = {vx, vy, vth};
//      mode 0: vref (3 x count twists), mode 1: one reference trajectory xt_ref (3 x ncols) for every instance
int ref_dwa_control_many(const signed char* data, unsigned xsize, unsigned ysize, double res, double xmin, double ymin,
                         const double* col, const double* cfg, const unsigned* samples, const double* x0,
                         const double* vb, int count, int mode, const double* ref, int ncols, double dt_ref,
                         int* found, double* u_opt)
{
  try
  {
    const ee::GridMap grid = occupancy_grid(data, xsize, ysize, res, xmin, ymin);
    const ee::Collision collision(col[0], col[1], col[2], col[3]);
    const ee::DynamicWindow dwa(collision, cfg[0], cfg[1], cfg[2], cfg[3], cfg[4], cfg[5], cfg[6], cfg[7], cfg[8],
                                cfg[9], cfg[10], cfg[11], samples[0], samples[1], samples[2]);
    mat xt_ref;
    if (mode == 1)
    {
      xt_ref.set_size(3, ncols);
      std::memcpy(xt_ref.memptr(), ref, sizeof(double) * 3 * ncols);
    }
    for (int i = 0; i < count; i++)
    {
      const auto r = mode == 0 ? dwa.control(grid, v3(x0 + 3 * i), v3(vb + 3 * i), v3(ref + 3 * i)) :
                                 dwa.control(grid, v3(x0 + 3 * i), v3(vb + 3 * i), xt_ref, dt_ref);
      found[i] = std::get<0>(r) ? 1 : 0;
      out_mat(std::get<1>(r), u_opt + 3 * i);
    }
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}
}  // extern "C"
