/*
 * ergodic_oracle.c -- TEST INFRASTRUCTURE ONLY (see ergodic_oracle.h).
 *
 * CPU restatement of the reference hot path in plain C doubles.  The
 * arithmetic follows the reference's association order where the reference
 * fixes one; Armadillo-internal summation order is taken
 * to be sequential.  Build with -ffp-contract=off so no FMA is introduced.
 *
 * Parity: pinned against the reference's known-answer vectors and against
 * the unmodified reference sources compiled against the test shim
 * (oracle/_ref) -- see tests/test_oracle_kat.py and tests/test_oracle_vs_ref.py.
 */
#include "ergodic_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* numerics.hpp                                                             */
/* ------------------------------------------------------------------------ */

/* numerics.hpp:67-70 */
int eo_almost_equal(double d1, double d2, double eps) { return fabs(d1 - d2) < eps ? 1 : 0; }

/* numerics.hpp:77-89 */
double eo_normalize_angle_pi(double rad)
{
  const double q = floor((rad + EO_PI) / (2.0 * EO_PI));
  rad = (rad + EO_PI) - q * 2.0 * EO_PI;
  if (rad < 0.0) rad += 2.0 * EO_PI;
  return rad - EO_PI;
}

/* numerics.hpp:273-298 (transform2d(angle) :243-249 applied to dqb) */
void eo_integrate_twist(const double x[3], const double u[3], double dt, double out[3])
{
  double dqb[3];
  if (eo_almost_equal(u[2], 0.0, 1.0e-12)) {
    dqb[0] = u[0] * dt;
    dqb[1] = u[1] * dt;
    dqb[2] = 0.0;
  } else {
    const double vb0 = u[0] * dt, vb1 = u[1] * dt, vb2 = u[2] * dt;
    dqb[0] = (vb0 * sin(vb2) + vb1 * (cos(vb2) - 1.0)) / vb2;
    dqb[1] = (vb1 * sin(vb2) + vb0 * (1.0 - cos(vb2))) / vb2;
    dqb[2] = vb2;
  }
  const double c = cos(x[2]), s = sin(x[2]);
  /* 3x3 mat-vec, row sums in column order */
  out[0] = x[0] + ((c * dqb[0] + (-s) * dqb[1]) + 0.0 * dqb[2]);
  out[1] = x[1] + ((s * dqb[0] + c * dqb[1]) + 0.0 * dqb[2]);
  out[2] = x[2] + ((0.0 * dqb[0] + 0.0 * dqb[1]) + 1.0 * dqb[2]);
}

/* ------------------------------------------------------------------------ */
/* occupancy-grid collision checking                                        */
/* ------------------------------------------------------------------------ */

/* Collision::checkCell collision.cpp:217-244 without the closest-obstacle book-keeping
 * (cfg.sqrd_obs, dx, dy feed minDistance / minDirection only).  cj, ci are `unsigned int`
 * in the reference: negative coordinates wrap and fail GridMap::gridBounds (grid.cpp:96-100). */
static int check_cell(const eo_grid *g, const eo_collision *c, int cx, int cy, int r_col,
                      unsigned int cj, unsigned int ci)
{
  if (!(ci <= g->ysize - 1u && cj <= g->xsize - 1u)) return 0;
  /* GridMap::getCell grid.cpp:177-184: int8 cell / 100.0 */
  const double cell = (double)g->data[(size_t)ci * g->xsize + cj] / 100.0;
  if (cell < c->occupied_threshold) return 0;
  const unsigned int dj = (unsigned int)cx - cj, di = (unsigned int)cy - ci;
  const int sqrd_obs = (int)(dj * dj + di * di);
  return sqrd_obs <= r_col * r_col;
}

/* Collision::bresenhamCircle collision.cpp:169-215 */
static int bresenham_circle(const eo_grid *g, const eo_collision *c, int cx, int cy, int r_col, int r)
{
  int x = -r, y = 0, err = 2 - 2 * r;
  while (x < 0) {
    if (check_cell(g, c, cx, cy, r_col, (unsigned int)(cx - x), (unsigned int)(cy + y))) return 1;
    if (check_cell(g, c, cx, cy, r_col, (unsigned int)(cx - y), (unsigned int)(cy - x))) return 1;
    if (check_cell(g, c, cx, cy, r_col, (unsigned int)(cx + x), (unsigned int)(cy - y))) return 1;
    if (check_cell(g, c, cx, cy, r_col, (unsigned int)(cx + y), (unsigned int)(cy + x))) return 1;
    r = err;
    if (r <= y) {
      y++;
      err += 2 * y + 1;
    }
    if (r > x || err > y) {
      x++;
      err += 2 * x + 1;
    }
  }
  return 0;
}

int eo_collision_check(const eo_grid *g, const eo_collision *c, const double pose[3])
{
  /* GridMap::world2Grid grid.cpp:143-160.  The reference casts floor() to unsigned int and
   * then to int (CollisionConfig); for a pose left of / below the map that is a wrap-around
   * on x86-64, i.e. the plain signed floor, which is what is restated here. */
  long long jl = (long long)floor((pose[0] - g->xmin) / g->resolution);
  long long il = (long long)floor((pose[1] - g->ymin) / g->resolution);
  unsigned int j = (unsigned int)jl, i = (unsigned int)il;
  if (j == g->xsize) j--;
  if (i == g->ysize) i--;
  const int cx = (int)j, cy = (int)i;
  /* collision.cpp:130-133 */
  const int r_bnd = (int)floor(c->boundary_radius / g->resolution);
  const int r_col = (int)floor((c->boundary_radius + c->obstacle_threshold) / g->resolution);
  const int r_max = (int)floor(c->search_radius / g->resolution);
  /* Collision::search collision.cpp:150-167 */
  for (int r = r_bnd; r <= r_max; r++)
    if (bresenham_circle(g, c, cx, cy, r_col, r)) return 1;
  return 0;
}

int eo_validate_control(const eo_grid *g, const eo_collision *c, const double x0[3],
                        const double u[3], double dt, double horizon)
{
  double x[3] = { x0[0], x0[1], x0[2] };
  const unsigned int steps = (unsigned int)fabs(horizon / dt);
  for (unsigned int i = 0; i < steps; i++) {
    double xn[3];
    eo_integrate_twist(x, u, dt, xn);
    xn[2] = eo_normalize_angle_pi(xn[2]);
    memcpy(x, xn, sizeof(x));
    if (eo_collision_check(g, c, x)) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------ */
/* DynamicWindow                                                            */
/* ------------------------------------------------------------------------ */

/* DynamicWindow::window dynamic_window.cpp:189-235 */
static void dwa_window(const eo_dwa *d, const double vb[3], unsigned int n[3], double lower[3], double delta[3])
{
  n[0] = d->vx_samples ? d->vx_samples : 1u;
  n[1] = d->vy_samples ? d->vy_samples : 1u;
  n[2] = d->vth_samples ? d->vth_samples : 1u;
  const double acc[3] = { d->acc_lim_x, d->acc_lim_y, d->acc_lim_th };
  const double vmin[3] = { d->min_vel_x, d->min_vel_y, d->min_rot_vel };
  const double vmax[3] = { d->max_vel_x, d->max_vel_y, d->max_rot_vel };
  for (int a = 0; a < 3; a++) {
    lower[a] = fmax(vb[a] - acc[a] * d->acc_dt, vmin[a]);
    const double upper = fmin(vb[a] + acc[a] * d->acc_dt, vmax[a]);
    delta[a] = n[a] > 1 ? (upper - lower[a]) / (double)(n[a] - 1) : 0.0;
  }
}

/* the two objectives (:237-257, :259-286); DBL_MAX on collision */
static double dwa_objective(const eo_grid *g, const eo_collision *c, const eo_dwa *d, const double x0[3],
                            const double u[3], const double *vref, const double *xt_ref, int ncols, double tf)
{
  double pose[3] = { x0[0], x0[1], x0[2] };
  const unsigned int steps = (unsigned int)fabs(d->horizon / d->dt);
  double t = 0.0, cost = 0.0;
  for (unsigned int i = 0; i < steps; i++) {
    double pn[3];
    eo_integrate_twist(pose, u, d->dt, pn);
    pn[2] = eo_normalize_angle_pi(pn[2]);
    memcpy(pose, pn, sizeof(pose));
    if (eo_collision_check(g, c, pose)) return DBL_MAX;
    if (xt_ref) {
      const unsigned int j = (unsigned int)round((double)(ncols - 1) * t / tf);
      const double dx = xt_ref[3 * j + 0] - pose[0], dy = xt_ref[3 * j + 1] - pose[1];
      cost += sqrt(dx * dx + dy * dy);
      cost += fabs(eo_normalize_angle_pi(eo_normalize_angle_pi(xt_ref[3 * j + 2]) - pose[2]));
      t += d->dt;
    }
  }
  if (xt_ref) return cost;
  const double e0 = vref[0] - u[0], e1 = vref[1] - u[1], e2 = vref[2] - u[2];
  return (e0 * e0 + e1 * e1) + e2 * e2;
}

static int dwa_search(const eo_grid *g, const eo_collision *c, const eo_dwa *d, const double x0[3],
                      const double vb[3], const double *vref, const double *xt_ref, int ncols, double tf,
                      double u_opt[3], double *min_cost_out)
{
  unsigned int n[3];
  double lower[3], delta[3];
  dwa_window(d, vb, n, lower, delta);
  double min_cost = DBL_MAX;
  u_opt[0] = u_opt[1] = u_opt[2] = 0.0;
  double vx = lower[0];
  for (unsigned int i = 0; i < n[0]; i++) {
    double vy = lower[1];
    for (unsigned int j = 0; j < n[1]; j++) {
      double w = lower[2];
      for (unsigned int k = 0; k < n[2]; k++) {
        const double u[3] = { vx, vy, w };
        const double cost = dwa_objective(g, c, d, x0, u, vref, xt_ref, ncols, tf);
        if (cost < min_cost) {
          min_cost = cost;
          memcpy(u_opt, u, sizeof(u));
        }
        w += delta[2];
      }
      vy += delta[1];
    }
    vx += delta[0];
  }
  if (min_cost_out) *min_cost_out = min_cost;
  return eo_almost_equal(min_cost, DBL_MAX, 1.0e-12) ? 0 : 1;
}

int eo_dwa_control_twist(const eo_grid *g, const eo_collision *c, const eo_dwa *d, const double x0[3],
                         const double vb[3], const double vref[3], double u_opt[3], double *min_cost)
{
  return dwa_search(g, c, d, x0, vb, vref, NULL, 0, 0.0, u_opt, min_cost);
}

int eo_dwa_control_traj(const eo_grid *g, const eo_collision *c, const eo_dwa *d, const double x0[3],
                        const double vb[3], const double *xt_ref, int ncols, double dt_ref,
                        double u_opt[3], double *min_cost)
{
  const double tf = (double)ncols * dt_ref; /* :148 */
  return dwa_search(g, c, d, x0, vb, NULL, xt_ref, ncols, tf, u_opt, min_cost);
}

/* ------------------------------------------------------------------------ */
/* models                                                                   */
/* ------------------------------------------------------------------------ */

/* SimpleCart cart.hpp:165-173, Omni omni.hpp:177-184 */
int eo_model_f(int model, const double x[3], const double u[3], double xdot[3])
{
  if (model == EO_MODEL_SIMPLE_CART) {
    if (!eo_almost_equal(u[1], 0.0, 1.0e-12)) return -1; /* cart.hpp:167-170 throws */
    xdot[0] = u[0] * cos(x[2]);
    xdot[1] = u[0] * sin(x[2]);
    xdot[2] = u[2];
    return 0;
  }
  xdot[0] = u[0] * cos(x[2]) - u[1] * sin(x[2]);
  xdot[1] = u[0] * sin(x[2]) + u[1] * cos(x[2]);
  xdot[2] = u[2];
  return 0;
}

/* SimpleCart cart.hpp:181-187, Omni omni.hpp:192-198.  A is 3x3 column-major. */
void eo_model_fdx(int model, const double x[3], const double u[3], double A[9])
{
  memset(A, 0, 9 * sizeof(double));
  if (model == EO_MODEL_SIMPLE_CART) {
    A[0 + 3 * 2] = -u[0] * sin(x[2]);
    A[1 + 3 * 2] = u[0] * cos(x[2]);
  } else {
    A[0 + 3 * 2] = -u[0] * sin(x[2]) - u[1] * cos(x[2]);
    A[1 + 3 * 2] = u[0] * cos(x[2]) - u[1] * sin(x[2]);
  }
}

/* SimpleCart cart.hpp:194-203, Omni omni.hpp:205-212.  B is 3x3 column-major. */
void eo_model_fdu(int model, const double x[3], double B[9])
{
  memset(B, 0, 9 * sizeof(double));
  if (model == EO_MODEL_SIMPLE_CART) {
    B[0 + 3 * 0] = cos(x[2]);
    B[1 + 3 * 0] = sin(x[2]);
    B[2 + 3 * 2] = 1.0;
  } else {
    B[0 + 3 * 0] = cos(x[2]);
    B[0 + 3 * 1] = -sin(x[2]);
    B[1 + 3 * 0] = sin(x[2]);
    B[1 + 3 * 1] = cos(x[2]);
    B[2 + 3 * 2] = 1.0;
  }
}

/* cart.hpp:93-101 */
void eo_cart_f(double wheel_radius, double wheel_base, const double x[3], const double u[2],
               double xdot[3])
{
  const double f = wheel_radius / 2.0;
  xdot[0] = f * ((u[0] + u[1]) * cos(x[2]));
  xdot[1] = f * ((u[0] + u[1]) * sin(x[2]));
  xdot[2] = f * ((u[1] - u[0]) / wheel_base);
}

/* cart.hpp:109-120 */
void eo_cart_fdx(double wheel_radius, double wheel_base, const double x[3], const double u[2],
                 double A[9])
{
  (void)wheel_base;
  memset(A, 0, 9 * sizeof(double));
  A[0 + 3 * 2] = -(wheel_radius / 2.0) * (u[0] + u[1]) * sin(x[2]);
  A[1 + 3 * 2] = (wheel_radius / 2.0) * (u[0] + u[1]) * cos(x[2]);
}

/* cart.hpp:127-141.  B is 3x2 column-major. */
void eo_cart_fdu(double wheel_radius, double wheel_base, const double x[3], double B[6])
{
  const double f = wheel_radius / 2.0;
  B[0 + 3 * 0] = f * cos(x[2]);
  B[0 + 3 * 1] = f * cos(x[2]);
  B[1 + 3 * 0] = f * sin(x[2]);
  B[1 + 3 * 1] = f * sin(x[2]);
  B[2 + 3 * 0] = f * (-1.0 / wheel_base);
  B[2 + 3 * 1] = f * (1.0 / wheel_base);
}

/* cart.hpp:79-85 */
void eo_cart_wheels2twist(double wheel_radius, double wheel_base, const double u[2],
                          double vb[3])
{
  vb[0] = wheel_radius / 2.0 * (u[0] + u[1]);
  vb[1] = 0.0;
  vb[2] = wheel_radius / (2.0 * wheel_base) * (u[1] - u[0]);
}

/* omni.hpp:98-110 */
void eo_mecanum_f(double r, double bx, double by, const double x[3], const double u[4],
                  double xdot[3])
{
  const double s = (r / 4.0) * sin(x[2]);
  const double c = (r / 4.0) * cos(x[2]);
  const double l = r / (4.0 * (bx + by));
  xdot[0] = u[0] * (s + c) + u[1] * (-s + c) + u[2] * (s + c) + u[3] * (-s + c);
  xdot[1] = u[0] * (s - c) + u[1] * (s + c) + u[2] * (s - c) + u[3] * (s + c);
  xdot[2] = -u[0] * l + u[1] * l + u[2] * l - u[3] * l;
}

/* omni.hpp:118-134 */
void eo_mecanum_fdx(double r, double bx, double by, const double x[3], const double u[4],
                    double A[9])
{
  (void)bx;
  (void)by;
  const double s = (r / 4.0) * sin(x[2]);
  const double c = (r / 4.0) * cos(x[2]);
  memset(A, 0, 9 * sizeof(double));
  A[0 + 3 * 2] = u[0] * (-s + c) + u[1] * (-s - c) + u[2] * (-s + c) + u[3] * (-s - c);
  A[1 + 3 * 2] = u[0] * (s + c) + u[1] * (-s + c) + u[2] * (s + c) + u[3] * (-s + c);
}

/* omni.hpp:141-151.  B is 3x4 column-major. */
void eo_mecanum_fdu(double r, double bx, double by, const double x[3], double B[12])
{
  const double s = (r / 4.0) * sin(x[2]);
  const double c = (r / 4.0) * cos(x[2]);
  const double l = r / (4.0 * (bx + by));
  const double row0[4] = { s + c, -s + c, s + c, -s + c };
  const double row1[4] = { s - c, s + c, s - c, s + c };
  const double row2[4] = { -l, l, l, -l };
  for (int j = 0; j < 4; j++) {
    B[0 + 3 * j] = row0[j];
    B[1 + 3 * j] = row1[j];
    B[2 + 3 * j] = row2[j];
  }
}

/* omni.hpp:80-90 */
void eo_mecanum_wheels2twist(double r, double bx, double by, const double u[4], double vb[3])
{
  const double l = 1.0 / (bx + by);
  const double Hp[3][4] = { { 1.0, 1.0, 1.0, 1.0 }, { -1.0, 1.0, -1.0, 1.0 }, { -l, l, l, -l } };
  for (int i = 0; i < 3; i++) {
    double acc = 0.0;
    for (int j = 0; j < 4; j++) acc += ((r / 4.0) * Hp[i][j]) * u[j];
    vb[i] = acc;
  }
}

/* ------------------------------------------------------------------------ */
/* integrator.hpp                                                           */
/* ------------------------------------------------------------------------ */

typedef int (*eo_dyn_fn)(const void *ctx, const double x[3], const double *u, double xdot[3]);

/* RungeKutta::step fwd integrator.hpp:176-184 */
static int rk4_step(eo_dyn_fn f, const void *ctx, double dt, const double x[3], const double *u,
                    double xn[3])
{
  double k1[3], k2[3], k3[3], k4[3], xs[3];
  if (f(ctx, x, u, k1)) return -1;
  for (int r = 0; r < 3; r++) xs[r] = x[r] + dt * (0.5 * k1[r]);
  if (f(ctx, xs, u, k2)) return -1;
  for (int r = 0; r < 3; r++) xs[r] = x[r] + dt * (0.5 * k2[r]);
  if (f(ctx, xs, u, k3)) return -1;
  for (int r = 0; r < 3; r++) xs[r] = x[r] + dt * k3[r];
  if (f(ctx, xs, u, k4)) return -1;
  for (int r = 0; r < 3; r++)
    xn[r] = x[r] + (dt / 6.0) * (((k1[r] + 2.0 * k2[r]) + 2.0 * k3[r]) + k4[r]);
  return 0;
}

static int dyn_twist(const void *ctx, const double x[3], const double *u, double xdot[3])
{
  return eo_model_f(*(const int *)ctx, x, u, xdot);
}

struct cart_ctx {
  double r, b;
};
static int dyn_cart(const void *ctx, const double x[3], const double *u, double xdot[3])
{
  const struct cart_ctx *c = (const struct cart_ctx *)ctx;
  eo_cart_f(c->r, c->b, x, u, xdot);
  return 0;
}

/* RungeKutta::solve fwd integrator.hpp:135-152 */
static int rk4_forward(eo_dyn_fn f, const void *ctx, int nu, double dt, double horizon,
                       const double x0[3], const double *ut, double *xt)
{
  const int steps = (int)(unsigned int)fabs(horizon / dt);
  double x[3] = { x0[0], x0[1], x0[2] };
  for (int i = 0; i < steps; i++) {
    double xn[3];
    if (rk4_step(f, ctx, dt, x, ut + (size_t)nu * i, xn)) return -1;
    xn[2] = eo_normalize_angle_pi(xn[2]);
    memcpy(x, xn, sizeof(x));
    memcpy(xt + 3 * (size_t)i, x, sizeof(x));
  }
  return steps;
}

int eo_rk4_forward(int model, double dt, double horizon, const double x0[3], const double *ut,
                   double *xt)
{
  return rk4_forward(dyn_twist, &model, 3, dt, horizon, x0, ut, xt);
}

int eo_rk4_forward_cart(double wheel_radius, double wheel_base, double dt, double horizon,
                        const double x0[3], const double *ut, double *xt)
{
  struct cart_ctx c = { wheel_radius, wheel_base };
  return rk4_forward(dyn_cart, &c, 2, dt, horizon, x0, ut, xt);
}

struct mecanum_ctx {
  double r, bx, by;
};
static int dyn_mecanum(const void *ctx, const double x[3], const double *u, double xdot[3])
{
  const struct mecanum_ctx *c = (const struct mecanum_ctx *)ctx;
  eo_mecanum_f(c->r, c->bx, c->by, x, u, xdot);
  return 0;
}

/* RungeKutta::solve fwd for models::Mecanum (omni.hpp:59-157), 4 wheel velocities per step */
int eo_rk4_forward_mecanum(double r, double bx, double by, double dt, double horizon, const double x0[3],
                           const double *ut, double *xt)
{
  struct mecanum_ctx c = { r, bx, by };
  return rk4_forward(dyn_mecanum, &c, 4, dt, horizon, x0, ut, xt);
}

/* entropy of one cell, numerics.hpp:164-179 */
double eo_entropy(double p)
{
  if (eo_almost_equal(0.0, p, 1.0e-12) || eo_almost_equal(1.0, p, 1.0e-12)) return 1e-3;
  else if (p < 0.0) return 0.7;
  return -p * log(p) - (1.0 - p) * log(1.0 - p);
}

/* entropy of every cell of an occupancy grid: p = GridMap::getCell = int8 / 100 (grid.cpp:177-184) */
void eo_entropy_grid(const signed char *cells, long long n, double *out)
{
  for (long long i = 0; i < n; i++) out[i] = eo_entropy((double)cells[i] / 100.0);
}

/* test tooling (tools/edge_study.py): `steps` constant-twist steps of integrate_twist + normalize_angle_PI
 * (numerics.hpp:273-298, 77-89, as validate_control / DynamicWindow chain them) for n poses; out: [steps][n][3] */
void eo_integrate_twist_chain(const double *x0, const double *u, double dt, long long n, int steps, double *out)
{
  for (long long i = 0; i < n; i++) {
    double x[3] = { x0[3 * i], x0[3 * i + 1], x0[3 * i + 2] };
    for (int k = 0; k < steps; k++) {
      double xn[3];
      eo_integrate_twist(x, u + 3 * i, dt, xn);
      xn[2] = eo_normalize_angle_pi(xn[2]);
      memcpy(x, xn, sizeof(x));
      memcpy(out + ((size_t)k * (size_t)n + (size_t)i) * 3, x, sizeof(x));
    }
  }
}

/* rhodot ergodic_control.hpp:65-69: -gdx - dbar - fdx.t()*rho */
static void rhodot(const double rho[3], const double gdx[3], const double dbar[3],
                   const double A[9], double out[3])
{
  for (int j = 0; j < 3; j++) {
    const double atr = (A[0 + 3 * j] * rho[0] + A[1 + 3 * j] * rho[1]) + A[2 + 3 * j] * rho[2];
    out[j] = (-gdx[j] - dbar[j]) - atr;
  }
}

/* RungeKutta::solve bwd integrator.hpp:154-174, step :186-194 */
void eo_rk4_backward(int model, double dt, int steps, const double rhoT[3], const double *xt,
                     const double *ut, const double *edx, const double *bdx, double *rhot)
{
  double rho[3] = { rhoT[0], rhoT[1], rhoT[2] };
  for (int i = steps; i-- > 0;) {
    double A[9], k1[3], k2[3], k3[3], k4[3], rs[3];
    const double *g = edx + 3 * (size_t)i, *b = bdx + 3 * (size_t)i;
    eo_model_fdx(model, xt + 3 * (size_t)i, ut + 3 * (size_t)i, A);
    rhodot(rho, g, b, A, k1);
    for (int r = 0; r < 3; r++) rs[r] = rho[r] - dt * (0.5 * k1[r]);
    rhodot(rs, g, b, A, k2);
    for (int r = 0; r < 3; r++) rs[r] = rho[r] - dt * (0.5 * k2[r]);
    rhodot(rs, g, b, A, k3);
    for (int r = 0; r < 3; r++) rs[r] = rho[r] - dt * k3[r];
    rhodot(rs, g, b, A, k4);
    for (int r = 0; r < 3; r++)
      rho[r] = rho[r] - dt / 6.0 * (((k1[r] + 2.0 * k2[r]) + 2.0 * k3[r]) + k4[r]);
    memcpy(rhot + 3 * (size_t)i, rho, sizeof(rho));
  }
}

/* ------------------------------------------------------------------------ */
/* basis.cpp                                                                */
/* ------------------------------------------------------------------------ */

/* Basis::Basis basis.cpp:48-77: k_(0,col)=j (kx), k_(1,col)=i (ky), col=i*nb+j */
void eo_basis_tables(int nb, long long *k, double *lamdak)
{
  int col = 0;
  for (int i = 0; i < nb; i++)
    for (int j = 0; j < nb; j++) {
      k[0 + 2 * col] = j;
      k[1 + 2 * col] = i;
      col++;
    }
  for (int i = 0; i < nb * nb; i++) {
    const long long ss = k[0 + 2 * i] * k[0 + 2 * i] + k[1 + 2 * i] * k[1 + 2 * i];
    lamdak[i] = 1.0 / pow((1.0 + sqrt((double)ss)), 1.5);
  }
}

/* Basis::fourierBasis basis.cpp:79-89 */
void eo_fourier_basis(double lx, double ly, int nb, const double x[2], double *fk)
{
  int col = 0;
  for (int i = 0; i < nb; i++)
    for (int j = 0; j < nb; j++) {
      fk[col] = cos((double)j * (EO_PI / lx) * x[0]) * cos((double)i * (EO_PI / ly) * x[1]);
      col++;
    }
}

/* Basis::gradFourierBasis basis.cpp:91-107; dfk is 2 x K column-major */
void eo_grad_fourier_basis(double lx, double ly, int nb, const double x[2], double *dfk)
{
  int col = 0;
  for (int i = 0; i < nb; i++)
    for (int j = 0; j < nb; j++) {
      const double k1 = (double)j * (EO_PI / lx);
      const double k2 = (double)i * (EO_PI / ly);
      dfk[0 + 2 * col] = -k1 * sin(k1 * x[0]) * cos(k2 * x[1]);
      dfk[1 + 2 * col] = -k2 * cos(k1 * x[0]) * sin(k2 * x[1]);
      col++;
    }
}

/* Basis::trajCoeff basis.cpp:109-120 */
void eo_traj_coeff(double lx, double ly, int nb, const double *xt, int ld, int ncols, double *ck)
{
  const int K = nb * nb;
  double *fk = (double *)malloc(sizeof(double) * (size_t)K);
  for (int k = 0; k < K; k++) ck[k] = 0.0;
  for (int c = 0; c < ncols; c++) {
    eo_fourier_basis(lx, ly, nb, xt + (size_t)ld * c, fk);
    for (int k = 0; k < K; k++) ck[k] += fk[k];
  }
  const double inv = 1.0 / (double)ncols;
  for (int k = 0; k < K; k++) ck[k] = inv * ck[k];
  free(fk);
}

/* Basis::spatialCoeff basis.cpp:122-133 */
void eo_spatial_coeff(double lx, double ly, int nb, const double *phi_vals,
                      const double *phi_grid, long long G, double *phik)
{
  const int K = nb * nb;
  double *fk = (double *)malloc(sizeof(double) * (size_t)K);
  for (int k = 0; k < K; k++) phik[k] = 0.0;
  for (long long c = 0; c < G; c++) {
    eo_fourier_basis(lx, ly, nb, phi_grid + 2 * c, fk);
    for (int k = 0; k < K; k++) phik[k] += fk[k] * phi_vals[c];
  }
  free(fk);
}

/* ------------------------------------------------------------------------ */
/* target.hpp / target.cpp                                                  */
/* ------------------------------------------------------------------------ */

/* Gaussian ctor target.hpp:68-71: inv() of the 2x2 diag(sigma^2) by cofactors */
void eo_gaussian_cov_inv(const double sigma[2], double cov_inv[4])
{
  const double a = sigma[0] * sigma[0], d = sigma[1] * sigma[1], b = 0.0, c = 0.0;
  const double det = a * d - b * c;
  cov_inv[0] = d / det;
  cov_inv[1] = -c / det;
  cov_inv[2] = -b / det;
  cov_inv[3] = a / det;
}

/* Gaussian::operator()(pt, trans) target.hpp:91-102 */
static double gaussian_eval(const double mu[2], const double ci[4], const double pt[2],
                            const double trans[2])
{
  const double d0 = pt[0] - (mu[0] - trans[0]);
  const double d1 = pt[1] - (mu[1] - trans[1]);
  /* diff.t()*cov_inv (1x2), then dot with diff */
  const double r0 = d0 * ci[0] + d1 * ci[1];
  const double r1 = d0 * ci[2] + d1 * ci[3];
  return exp(-0.5 * (r0 * d0 + r1 * d1));
}

/* Target::evaluate target.cpp:68-76 */
double eo_target_evaluate(int ng, const double *mu, const double *sigma, const double pt[2],
                          const double trans[2])
{
  double val = 0.0;
  for (int g = 0; g < ng; g++) {
    double ci[4];
    eo_gaussian_cov_inv(sigma + 2 * g, ci);
    val += gaussian_eval(mu + 2 * g, ci, pt, trans);
  }
  return val;
}

/* Target::fill target.cpp:78-89 */
void eo_target_fill(int ng, const double *mu, const double *sigma, const double trans[2],
                    const double *phi_grid, long long G, double *phi_vals)
{
  double *ci = (double *)malloc(sizeof(double) * 4 * (size_t)(ng > 0 ? ng : 1));
  for (int g = 0; g < ng; g++) eo_gaussian_cov_inv(sigma + 2 * g, ci + 4 * g);
  double total = 0.0;
  for (long long c = 0; c < G; c++) {
    double val = 0.0;
    for (int g = 0; g < ng; g++) val += gaussian_eval(mu + 2 * g, ci + 4 * g, phi_grid + 2 * c, trans);
    phi_vals[c] = val;
    total += val;
  }
  for (long long c = 0; c < G; c++) phi_vals[c] /= total;
  free(ci);
}

/* ergodic_control.hpp:387-388 with grid.hpp:61-64 */
void eo_target_grid_dims(double lx, double ly, double resolution, int *nx, int *ny)
{
  *nx = (int)(unsigned int)round((lx - 0.0) / resolution) + 1;
  *ny = (int)(unsigned int)round((ly - 0.0) / resolution) + 1;
}

/* ergodic_control.hpp:391-408: accumulated coordinates, y outer, x inner */
void eo_target_grid(double resolution, int nx, int ny, double *phi_grid)
{
  long long col = 0;
  double y = 0.0;
  for (int i = 0; i < ny; i++) {
    double x = 0.0;
    for (int j = 0; j < nx; j++) {
      phi_grid[0 + 2 * col] = x;
      phi_grid[1 + 2 * col] = y;
      col++;
      x += resolution;
    }
    y += resolution;
  }
}

void eo_phik_from_grid(const double *phi, int nx, int ny, double resolution, double lx,
                       double ly, int nb, double *phik, double *phi_sum)
{
  const int K = nb * nb;
  double total = 0.0;
  for (long long c = 0; c < (long long)nx * ny; c++) total += phi[c];
  double *fk = (double *)malloc(sizeof(double) * (size_t)K);
  for (int k = 0; k < K; k++) phik[k] = 0.0;
  double y = 0.0;
  for (int i = 0; i < ny; i++) {
    double x = 0.0;
    for (int j = 0; j < nx; j++) {
      const double pt[2] = { x, y };
      const double v = phi[(size_t)i * nx + j] / total;
      eo_fourier_basis(lx, ly, nb, pt, fk);
      for (int k = 0; k < K; k++) phik[k] += fk[k] * v;
      x += resolution;
    }
    y += resolution;
  }
  if (phi_sum) *phi_sum = total;
  free(fk);
}

/* Row block [row_begin, row_begin + nrows) of eo_phik_from_grid for grids too large for one thread (C3: 8192^2 cells):
 * the SAME per-cell arithmetic -- F_k = cos(kx (PI / lx) x) cos(ky (PI / ly) y) with the accumulated coordinates,
 * acc_k += F_k * (phi / total) -- but the two cosine factors are tabulated per column / per row instead of being
 * re-evaluated in every cell: the same inputs give the same doubles, only 2K libm calls per cell are saved.  `total`
 * is sum(phi) over the WHOLE grid (target.cpp:87).  Partial sums of different row blocks are added by the caller. */
void eo_phik_rows(const double *phi_rows, int nx, int row_begin, int nrows, double resolution, double lx, double ly,
                  int nb, double total, double *acc /* nb*nb, overwritten */)
{
  const int K = nb * nb;
  double *cxt = (double *)malloc(sizeof(double) * (size_t)nx * nb);
  double *cyr = (double *)malloc(sizeof(double) * (size_t)nb);
  double *fk = (double *)malloc(sizeof(double) * (size_t)K);
  for (int k = 0; k < K; k++) acc[k] = 0.0;
  double x = 0.0;
  for (int j = 0; j < nx; j++) {
    for (int kx = 0; kx < nb; kx++) cxt[(size_t)j * nb + kx] = cos(((double)kx * (EO_PI / lx)) * x); /* basis.cpp:85 */
    x += resolution;
  }
  double y = 0.0;
  for (int i = 0; i < row_begin; i++) y += resolution; /* the accumulated y of configTarget (:394-407) */
  for (int i = 0; i < nrows; i++) {
    for (int ky = 0; ky < nb; ky++) cyr[ky] = cos(((double)ky * (EO_PI / ly)) * y);
    for (int j = 0; j < nx; j++) {
      const double v = phi_rows[(size_t)i * nx + j] / total;
      const double *cxj = cxt + (size_t)j * nb;
      for (int ky = 0; ky < nb; ky++)
        for (int kx = 0; kx < nb; kx++) fk[ky * nb + kx] = cxj[kx] * cyr[ky];
      for (int k = 0; k < K; k++) acc[k] += fk[k] * v;
    }
    y += resolution;
  }
  free(cxt);
  free(cyr);
  free(fk);
}

/* ------------------------------------------------------------------------ */
/* ErgodicControl                                                           */
/* ------------------------------------------------------------------------ */

struct eo_controller {
  int model;
  double dt, horizon, resolution, expl_weight;
  int steps, nb, K;
  long long buffer_size;
  int batch_size;
  double Rinv[9], umin[3], umax[3];
  double *ut;   /* 3 x steps */
  double *phik; /* K */
  double *lamdak;
  double rhoT[3];
  double map_pos[2];
  double pose[3];
  double lx, ly; /* basis_.lx_, basis_.ly_ */
  /* replay buffer: insertion-ordered states (keys are insertion indices) */
  double *mem;
  long long mem_size, mem_cap;
  /* target */
  int ng;
  double *mu, *sigma;
  /* by-products of the last control() */
  double *ck, *edx, *bdx, *rhot, *xtf;
  double metric;
};

/* ctor ergodic_control.hpp:188-222 */
eo_controller *eo_create(int model, double dt, double horizon, double resolution,
                         double expl_weight, int num_basis, long long buffer_size,
                         int batch_size, const double Rinv[9], const double umin[3],
                         const double umax[3])
{
  const int steps = (int)(unsigned int)fabs(horizon / dt);
  if (steps == 1) return NULL; /* :212-216 throws std::invalid_argument */
  eo_controller *c = (eo_controller *)calloc(1, sizeof(*c));
  c->model = model;
  c->dt = dt;
  c->horizon = horizon;
  c->resolution = resolution;
  c->expl_weight = expl_weight;
  c->steps = steps;
  c->nb = num_basis;
  c->K = num_basis * num_basis;
  c->buffer_size = buffer_size;
  c->batch_size = batch_size;
  memcpy(c->Rinv, Rinv, sizeof(c->Rinv));
  memcpy(c->umin, umin, sizeof(c->umin));
  memcpy(c->umax, umax, sizeof(c->umax));
  c->ut = (double *)calloc((size_t)3 * (steps > 0 ? steps : 1), sizeof(double));
  c->phik = (double *)calloc((size_t)c->K + 1, sizeof(double));
  c->lamdak = (double *)calloc((size_t)c->K + 1, sizeof(double));
  long long *k = (long long *)malloc(sizeof(long long) * 2 * ((size_t)c->K + 1));
  eo_basis_tables(c->nb, k, c->lamdak);
  free(k);
  c->lx = 0.0; /* Basis(0.0, 0.0, num_basis) :208 */
  c->ly = 0.0;
  c->ck = (double *)calloc((size_t)c->K + 1, sizeof(double));
  c->edx = (double *)calloc((size_t)3 * (steps > 0 ? steps : 1), sizeof(double));
  c->bdx = (double *)calloc((size_t)3 * (steps > 0 ? steps : 1), sizeof(double));
  c->rhot = (double *)calloc((size_t)3 * (steps > 0 ? steps : 1), sizeof(double));
  c->xtf = (double *)calloc((size_t)3 * (steps > 0 ? steps : 1), sizeof(double));
  return c;
}

void eo_destroy(eo_controller *c)
{
  if (!c) return;
  free(c->ut);
  free(c->phik);
  free(c->lamdak);
  free(c->mem);
  free(c->mu);
  free(c->sigma);
  free(c->ck);
  free(c->edx);
  free(c->bdx);
  free(c->rhot);
  free(c->xtf);
  free(c);
}

/* setTarget :357-360 */
void eo_set_target(eo_controller *c, int ng, const double *mu, const double *sigma)
{
  free(c->mu);
  free(c->sigma);
  c->ng = ng;
  c->mu = (double *)malloc(sizeof(double) * 2 * (size_t)(ng > 0 ? ng : 1));
  c->sigma = (double *)malloc(sizeof(double) * 2 * (size_t)(ng > 0 ? ng : 1));
  memcpy(c->mu, mu, sizeof(double) * 2 * (size_t)ng);
  memcpy(c->sigma, sigma, sizeof(double) * 2 * (size_t)ng);
}

/* configTarget :363-416 */
int eo_config_target(eo_controller *c, double xmin, double xmax, double ymin, double ymax)
{
  c->map_pos[0] = xmin;
  c->map_pos[1] = ymin;
  const double mx = xmax - xmin;
  const double my = ymax - ymin;
  if (eo_almost_equal(mx, c->lx, 1.0e-12) && eo_almost_equal(my, c->ly, 1.0e-12)) return 0;
  c->lx = mx;
  c->ly = my;
  int nx, ny;
  eo_target_grid_dims(c->lx, c->ly, c->resolution, &nx, &ny);
  const long long G = (long long)nx * ny;
  double *grid = (double *)malloc(sizeof(double) * 2 * (size_t)G);
  double *vals = (double *)malloc(sizeof(double) * (size_t)G);
  eo_target_grid(c->resolution, nx, ny, grid);
  eo_target_fill(c->ng, c->mu, c->sigma, c->map_pos, grid, G, vals);
  eo_spatial_coeff(c->lx, c->ly, c->nb, vals, grid, G, c->phik);
  free(grid);
  free(vals);
  return 1;
}

/* ReplayBuffer::append buffer.cpp:54-62 (silently drops when full) */
void eo_add_state_memory(eo_controller *c, const double x[3])
{
  if (c->mem_size >= c->buffer_size) return;
  if (c->mem_size == c->mem_cap) {
    c->mem_cap = c->mem_cap ? 2 * c->mem_cap : 64;
    c->mem = (double *)realloc(c->mem, sizeof(double) * 3 * (size_t)c->mem_cap);
  }
  memcpy(c->mem + 3 * c->mem_size, x, 3 * sizeof(double));
  c->mem_size++;
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (hi < v ? hi : v); }

/* control :225-311 */
int eo_control(eo_controller *c, double xmin, double xmax, double ymin, double ymax,
               const double x[3], const int *mem_idx, double u0[3])
{
  const int N = c->steps, K = c->K;
  memcpy(c->pose, x, sizeof(c->pose)); /* :227 */
  eo_config_target(c, xmin, xmax, ymin, ymax); /* :230 */

  /* :233-234 shift left by one column, zero the last */
  if (N >= 2) memmove(c->ut, c->ut + 3, sizeof(double) * 3 * (size_t)(N - 1));
  if (N >= 1) c->ut[3 * (N - 1) + 0] = c->ut[3 * (N - 1) + 1] = c->ut[3 * (N - 1) + 2] = 0.0;

  /* :237 forward simulation (map frame) */
  double *traj = (double *)malloc(sizeof(double) * 3 * (size_t)(N > 0 ? N : 1));
  if (eo_rk4_forward(c->model, c->dt, c->horizon, c->pose, c->ut, traj) < 0) {
    free(traj);
    return -1;
  }

  /* :240 ReplayBuffer::sampleMemory buffer.cpp:64-111 */
  long long M = 0;
  if (c->mem_size > 0) M = (c->mem_size <= c->batch_size) ? c->mem_size : c->batch_size;
  const int T = (int)M + N;
  double *xt_total = (double *)malloc(sizeof(double) * 3 * (size_t)(T > 0 ? T : 1));
  for (long long i = 0; i < M; i++) {
    const long long src = (c->mem_size <= c->batch_size) ? i : (long long)mem_idx[i];
    memcpy(xt_total + 3 * i, c->mem + 3 * src, 3 * sizeof(double));
  }
  memcpy(xt_total + 3 * M, traj, sizeof(double) * 3 * (size_t)N);

  /* :243-244 map frame -> fourier frame */
  for (int i = 0; i < T; i++) {
    xt_total[3 * i + 0] -= c->map_pos[0];
    xt_total[3 * i + 1] -= c->map_pos[1];
  }
  const double *xt = xt_total + 3 * M; /* :264 */
  memcpy(c->xtf, xt, sizeof(double) * 3 * (size_t)N);

  /* :267 */
  eo_traj_coeff(c->lx, c->ly, c->nb, xt_total, 3, T, c->ck);

  /* gradErgodicMetric :419-436 */
  double *fd = (double *)malloc(sizeof(double) * (size_t)K);
  double *dfk = (double *)malloc(sizeof(double) * 2 * (size_t)K);
  c->metric = 0.0;
  for (int k = 0; k < K; k++) {
    fd[k] = c->lamdak[k] * (c->ck[k] - c->phik[k]); /* :422 */
    c->metric += c->lamdak[k] * (c->ck[k] - c->phik[k]) * (c->ck[k] - c->phik[k]);
  }
  for (int i = 0; i < N; i++) {
    eo_grad_fourier_basis(c->lx, c->ly, c->nb, xt + 3 * i, dfk);
    double e0 = 0.0, e1 = 0.0;
    for (int k = 0; k < K; k++) {
      e0 += dfk[0 + 2 * k] * fd[k];
      e1 += dfk[1 + 2 * k] * fd[k];
    }
    c->edx[3 * i + 0] = e0 * c->expl_weight; /* :433 */
    c->edx[3 * i + 1] = e1 * c->expl_weight;
    c->edx[3 * i + 2] = 0.0; /* :430 */
  }
  free(fd);
  free(dfk);

  /* gradBarrier :454-474 */
  {
    const double weight = 25.0, eps = 0.05;
    for (int i = 0; i < N; i++) {
      const double px = xt[3 * i + 0], py = xt[3 * i + 1];
      double b0 = 0.0, b1 = 0.0;
      b0 += 2.0 * (double)(px > c->lx - eps) * (px - (c->lx - eps));
      b1 += 2.0 * (double)(py > c->ly - eps) * (py - (c->ly - eps));
      b0 += 2.0 * (double)(px < eps) * (px - eps);
      b1 += 2.0 * (double)(py < eps) * (py - eps);
      c->bdx[3 * i + 0] = b0 * weight;
      c->bdx[3 * i + 1] = b1 * weight;
      c->bdx[3 * i + 2] = 0.0;
    }
  }

  /* :277 backwards pass */
  eo_rk4_backward(c->model, c->dt, N, c->rhoT, xt, c->ut, c->edx, c->bdx, c->rhot);

  /* updateControl :439-451: u = -Rinv * B^T * rho (B^T rho first), clamp */
  for (int i = 0; i < N; i++) {
    double B[9], btr[3];
    const double *rho = c->rhot + 3 * i;
    eo_model_fdu(c->model, xt + 3 * i, B);
    for (int j = 0; j < 3; j++)
      btr[j] = (B[0 + 3 * j] * rho[0] + B[1 + 3 * j] * rho[1]) + B[2 + 3 * j] * rho[2];
    for (int r = 0; r < 3; r++) {
      const double v =
          -((c->Rinv[r + 3 * 0] * btr[0] + c->Rinv[r + 3 * 1] * btr[1]) + c->Rinv[r + 3 * 2] * btr[2]);
      c->ut[3 * i + r] = clampd(v, c->umin[r], c->umax[r]);
    }
  }

  free(traj);
  free(xt_total);
  if (N >= 1) memcpy(u0, c->ut, 3 * sizeof(double)); /* :310 */
  return 0;
}

/* optTraj :314-317 */
int eo_opt_traj(const eo_controller *c, double *xt)
{
  return eo_rk4_forward(c->model, c->dt, c->horizon, c->pose, c->ut, xt);
}

int eo_steps(const eo_controller *c) { return c->steps; }
int eo_num_coeff(const eo_controller *c) { return c->K; }
long long eo_memory_size(const eo_controller *c) { return c->mem_size; }
void eo_get_ut(const eo_controller *c, double *ut)
{
  memcpy(ut, c->ut, sizeof(double) * 3 * (size_t)c->steps);
}
void eo_set_ut(eo_controller *c, const double *ut)
{
  memcpy(c->ut, ut, sizeof(double) * 3 * (size_t)c->steps);
}
void eo_get_phik(const eo_controller *c, double *phik)
{
  memcpy(phik, c->phik, sizeof(double) * (size_t)c->K);
}
void eo_set_phik(eo_controller *c, const double *phik, double lx, double ly)
{
  memcpy(c->phik, phik, sizeof(double) * (size_t)c->K);
  c->lx = lx;
  c->ly = ly;
}
void eo_get_last(const eo_controller *c, double *ck, double *metric, double *edx, double *bdx,
                 double *rhot, double *xt_fourier)
{
  const size_t n3 = sizeof(double) * 3 * (size_t)c->steps;
  if (ck) memcpy(ck, c->ck, sizeof(double) * (size_t)c->K);
  if (metric) *metric = c->metric;
  if (edx) memcpy(edx, c->edx, n3);
  if (bdx) memcpy(bdx, c->bdx, n3);
  if (rhot) memcpy(rhot, c->rhot, n3);
  if (xt_fourier) memcpy(xt_fourier, c->xtf, n3);
}

int eo_control_many(eo_controller **cs, int count, double xmin, double xmax, double ymin,
                    double ymax, const double *x, double *u0)
{
  int rc = 0;
  for (int i = 0; i < count; i++)
    rc |= eo_control(cs[i], xmin, xmax, ymin, ymax, x + 3 * (size_t)i, NULL, u0 + 3 * (size_t)i);
  return rc;
}
