/*
 * ergodic_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded, IEEE-double restatement of the reference
 * (bostoncleek/ergodic_exploration) receding-horizon ergodic controller hot
 * path.  It exists to CHECK the CUDA path; nothing in the product
 * (ergodic_exploration_b200/, include/) may include, link or call it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it.
 *
 * Parity pinning: this restatement is pinned (a) against the reference's own
 * known-answer vectors (test/test_cart.cpp, test/test_omni.cpp,
 * test/test_integrator.cpp) and (b) against the UNMODIFIED reference sources
 * compiled in place against a test-only Armadillo/ROS header shim
 * (oracle/_ref, see oracle/Makefile); see tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference root).  Layouts are Armadillo column-major: a 3xN matrix
 * is N consecutive (x, y, theta) triples.
 */
#ifndef ERGODIC_ORACLE_H
#define ERGODIC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define EO_PI 3.14159265358979323846 /* numerics.hpp:58 */

enum { EO_MODEL_SIMPLE_CART = 0, EO_MODEL_OMNI = 1 };

/* ---- numerics.hpp ------------------------------------------------------ */
int eo_almost_equal(double d1, double d2, double eps);             /* :67-70  */
double eo_normalize_angle_pi(double rad);                          /* :77-89  */
void eo_integrate_twist(const double x[3], const double u[3], double dt,
                        double out[3]);                            /* :273-298 */

/* ---- models/cart.hpp, models/omni.hpp ---------------------------------- */
/* 3-twist models (the only ones ErgodicControl can run) */
int eo_model_f(int model, const double x[3], const double u[3], double xdot[3]);
void eo_model_fdx(int model, const double x[3], const double u[3], double A[9]);
void eo_model_fdu(int model, const double x[3], double B[9]);
/* wheel models (forward rollout only) */
void eo_cart_f(double wheel_radius, double wheel_base, const double x[3],
               const double u[2], double xdot[3]);                 /* cart.hpp:93-101 */
void eo_cart_fdx(double wheel_radius, double wheel_base, const double x[3],
                 const double u[2], double A[9]);                  /* cart.hpp:109-120 */
void eo_cart_fdu(double wheel_radius, double wheel_base, const double x[3],
                 double B[6]);                                     /* cart.hpp:127-141 */
void eo_cart_wheels2twist(double wheel_radius, double wheel_base,
                          const double u[2], double vb[3]);        /* cart.hpp:79-85 */
void eo_mecanum_f(double r, double bx, double by, const double x[3],
                  const double u[4], double xdot[3]);              /* omni.hpp:98-110 */
void eo_mecanum_fdx(double r, double bx, double by, const double x[3],
                    const double u[4], double A[9]);               /* omni.hpp:118-134 */
void eo_mecanum_fdu(double r, double bx, double by, const double x[3],
                    double B[12]);                                 /* omni.hpp:141-151 */
void eo_mecanum_wheels2twist(double r, double bx, double by, const double u[4],
                             double vb[3]);                        /* omni.hpp:80-90 */

/* ---- integrator.hpp ---------------------------------------------------- */
/* RungeKutta::solve fwd :135-152 (+step :176-184) for the 3-twist models.
 * xt is 3 x steps, steps = (unsigned)fabs(horizon/dt).  Returns steps or -1
 * when SimpleCart's u(1)!=0 guard (cart.hpp:167-170) would throw. */
int eo_rk4_forward(int model, double dt, double horizon, const double x0[3],
                   const double *ut, double *xt);
/* same, Cart wheel model (test/test_integrator.cpp:44-73), ut is 2 x steps */
int eo_rk4_forward_cart(double wheel_radius, double wheel_base, double dt,
                        double horizon, const double x0[3], const double *ut,
                        double *xt);
/* RungeKutta::solve bwd :154-174 (+step :186-194, rhodot ergodic_control.hpp:65-69) */
int eo_rk4_forward_mecanum(double r, double bx, double by, double dt, double horizon,
                           const double x0[3], const double *ut /*4 x steps*/, double *xt);
double eo_entropy(double p);                                       /* numerics.hpp:164-179 */
void eo_entropy_grid(const signed char *cells, long long n, double *out); /* + grid.cpp:177-184 */
void eo_integrate_twist_chain(const double *x0, const double *u, double dt, long long n, int steps, double *out);
void eo_rk4_backward(int model, double dt, int steps, const double rhoT[3],
                     const double *xt, const double *ut, const double *edx,
                     const double *bdx, double *rhot);

/* ---- basis.cpp --------------------------------------------------------- */
void eo_basis_tables(int nb, long long *k /*2 x nb^2*/, double *lamdak);       /* :48-77 */
void eo_fourier_basis(double lx, double ly, int nb, const double x[2], double *fk); /* :79-89 */
void eo_grad_fourier_basis(double lx, double ly, int nb, const double x[2],
                           double *dfk /*2 x nb^2*/);                          /* :91-107 */
/* xt has `ld` rows per column (2 or 3); only rows 0,1 are read          :109-120 */
void eo_traj_coeff(double lx, double ly, int nb, const double *xt, int ld, int ncols,
                   double *ck);
/* phi_grid 2 x G, phi_vals G                                            :122-133 */
void eo_spatial_coeff(double lx, double ly, int nb, const double *phi_vals,
                      const double *phi_grid, long long G, double *phik);

/* ---- target.hpp / target.cpp ------------------------------------------ */
/* Gaussian ctor target.hpp:68-71: cov = diag(sigma^2), cov_inv = inv(cov)
 * (2x2 cofactor inverse).  cov_inv is 2x2 column-major. */
void eo_gaussian_cov_inv(const double sigma[2], double cov_inv[4]);
/* Target::evaluate target.cpp:68-76 with Gaussian::operator()(pt,trans) target.hpp:91-102 */
double eo_target_evaluate(int ng, const double *mu, const double *sigma,
                          const double pt[2], const double trans[2]);
/* Target::fill target.cpp:78-89 */
void eo_target_fill(int ng, const double *mu, const double *sigma, const double trans[2],
                    const double *phi_grid, long long G, double *phi_vals);
/* the grid construction of configTarget ergodic_control.hpp:385-408 */
void eo_target_grid_dims(double lx, double ly, double resolution, int *nx, int *ny);
void eo_target_grid(double resolution, int nx, int ny, double *phi_grid /*2 x nx*ny*/);

/* Streaming restatement of the SAME naive arithmetic as Target::fill's
 * normalisation + Basis::spatialCoeff for an arbitrary (un-normalised) dense
 * density phi[ny][nx] (x fastest) on the configTarget grid (accumulated
 * coordinates): phik[k] = sum_c F_k(p_c) * (phi_c / sum phi).  No K x G
 * temporary.  Used for config C3 where the literal code cannot run. */
void eo_phik_from_grid(const double *phi, int nx, int ny, double resolution, double lx,
                       double ly, int nb, double *phik, double *phi_sum);
void eo_phik_rows(const double *phi_rows, int nx, int row_begin, int nrows, double resolution, double lx, double ly,
                  int nb, double total, double *acc);

/* ---- ErgodicControl (ergodic_control.hpp) ------------------------------ */
/* ---- occupancy-grid collision checking ------------- */
/* GridMap geometry (grid.hpp / grid.cpp): int8 cells, row-major (i = y row, j = x column) */
typedef struct eo_grid {
  const signed char *data; /* ysize x xsize */
  unsigned int xsize, ysize;
  double resolution, xmin, ymin;
} eo_grid;
/* Collision (collision.hpp:85-165, collision.cpp:46-64) */
typedef struct eo_collision {
  double boundary_radius, search_radius, obstacle_threshold, occupied_threshold;
} eo_collision;
/* Collision::collisionCheck collision.cpp:126-143 (+ search :150-167, bresenhamCircle
 * :169-215, checkCell :217-244, GridMap::world2Grid grid.cpp:143-160); 1 = collision */
int eo_collision_check(const eo_grid *g, const eo_collision *c, const double pose[3]);
/* validate_control numerics.hpp:312-330; 1 = collision free */
int eo_validate_control(const eo_grid *g, const eo_collision *c, const double x0[3],
                        const double u[3], double dt, double horizon);

/* ---- DynamicWindow (dynamic_window.hpp:60-172) ----- */
typedef struct eo_dwa {
  double dt, horizon, acc_dt, acc_lim_x, acc_lim_y, acc_lim_th;
  double max_vel_x, min_vel_x, max_vel_y, min_vel_y, max_rot_vel, min_rot_vel;
  unsigned int vx_samples, vy_samples, vth_samples; /* 0 is promoted to 1 (dynamic_window.cpp:71-91) */
} eo_dwa;
/* DynamicWindow::control(grid, x0, vb, vref) dynamic_window.cpp:93-139 (window :189-235,
 * objective :237-257); returns 1 if a collision-free twist exists; u_opt, min_cost out */
int eo_dwa_control_twist(const eo_grid *g, const eo_collision *c, const eo_dwa *d, const double x0[3],
                         const double vb[3], const double vref[3], double u_opt[3], double *min_cost);
/* DynamicWindow::control(grid, x0, vb, xt_ref, dt_ref) :141-187 (objective :259-286);
 * xt_ref is 3 x ncols column-major */
int eo_dwa_control_traj(const eo_grid *g, const eo_collision *c, const eo_dwa *d, const double x0[3],
                        const double vb[3], const double *xt_ref, int ncols, double dt_ref,
                        double u_opt[3], double *min_cost);

typedef struct eo_controller eo_controller;

/* ctor :188-222.  Rinv is 3x3 column-major.  Returns NULL when steps==1
 * (the reference throws std::invalid_argument :212-216). */
eo_controller *eo_create(int model, double dt, double horizon, double resolution,
                         double expl_weight, int num_basis, long long buffer_size,
                         int batch_size, const double Rinv[9], const double umin[3],
                         const double umax[3]);
void eo_destroy(eo_controller *c);
/* setTarget :357-360 */
void eo_set_target(eo_controller *c, int ng, const double *mu, const double *sigma);
/* configTarget :363-416 (grid bounds only).  Returns 1 if phik was rebuilt. */
int eo_config_target(eo_controller *c, double xmin, double xmax, double ymin, double ymax);
/* addStateMemory :345-348 / ReplayBuffer::append buffer.cpp:54-62 */
void eo_add_state_memory(eo_controller *c, const double x[3]);
/* control :225-311.  mem_idx: indices of the sampled past states when the
 * buffer holds more than batch_size entries (replaces arma::randi,
 * buffer.cpp:98 -- made an explicit input); ignored (may be
 * NULL) otherwise.  Returns 0, or -1 if SimpleCart's guard would throw. */
int eo_control(eo_controller *c, double xmin, double xmax, double ymin, double ymax,
               const double x[3], const int *mem_idx, double u0[3]);
/* optTraj :314-317 */
int eo_opt_traj(const eo_controller *c, double *xt);
/* state access (reference private members, exposed for teacher forcing) */
int eo_steps(const eo_controller *c);
int eo_num_coeff(const eo_controller *c);
long long eo_memory_size(const eo_controller *c);
void eo_get_ut(const eo_controller *c, double *ut);
void eo_set_ut(eo_controller *c, const double *ut);
void eo_get_phik(const eo_controller *c, double *phik);
void eo_set_phik(eo_controller *c, const double *phik, double lx, double ly);
/* by-products of the last control(): c_k, the ergodic metric
 * sum_k lamda_k (c_k - phi_k)^2 (not computed by the reference),
 * edx, bdx, rhot, and the Fourier-frame trajectory */
void eo_get_last(const eo_controller *c, double *ck, double *metric, double *edx,
                 double *bdx, double *rhot, double *xt_fourier);

/* Batched driver used for the CPU baseline: runs `count` independent
 * controllers (one eo_controller each, created by the caller) through one
 * control() call each, sequentially.  x is 3 x count, u0 is 3 x count. */
int eo_control_many(eo_controller **cs, int count, double xmin, double xmax, double ymin,
                    double ymax, const double *x, double *u0);

#ifdef __cplusplus
}
#endif
#endif
