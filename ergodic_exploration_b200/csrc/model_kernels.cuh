// model_kernels.cuh -- the four kinematic models and the forward RK4 integrator, batched (sm_100a, FP64).
//
// Replaces, per instance: models::Cart (models/cart.hpp:60-145), models::SimpleCart (:152-206),
// models::Mecanum (models/omni.hpp:59-157), models::Omni (:164-215) -- operator(), fdx, fdu, wheels2Twist --
// and RungeKutta::solve / step for the forward problem (integrator.hpp:135-152, 176-184; the heading is
// wrapped with normalize_angle_PI after every step, numerics.hpp:77-89).
//
// The fused control() kernel (solve_kernel.cuh) integrates the two 3-twist models with warp scans; this file is
// the general path: any model, any control dimension, ONE THREAD PER INSTANCE walking the horizon step by step
// in exactly the reference's expression order (explicitly rounded operations, no FMA contraction), so the only
// difference to the reference is the last ulp of sin / cos.  It backs the RungeKutta adapter class and the
// model adapters' operator() / fdx / fdu; it is also the independent check of the scan formulation.
// Layouts are Armadillo's: ut is [count][steps][nu] (column-major nu x steps per instance), xt [count][steps][3].
#pragma once

#include "common.cuh"
#include "solve_kernel.cuh"  // kModelSimpleCart, kModelOmni

namespace eb
{
constexpr int kModelCart = 2;     // 2 wheel velocities; params = { wheel_radius, wheel_base }
constexpr int kModelMecanum = 3;  // 4 wheel velocities; params = { wheel_radius, wheel_base_x, wheel_base_y }

__host__ __device__ inline int model_controls(int model)
{
  return model == kModelCart ? 2 : model == kModelMecanum ? 4 : 3;
}

struct ModelParams
{
  int model;
  double a, b, c;  // wheel_radius, wheel_base (_x), wheel_base_y
};

// f(x, u); returns false where the reference throws (SimpleCart with a y-velocity, cart.hpp:167-170)
__device__ __forceinline__ bool model_eval(const ModelParams& m, double th, const double* u, double& f0, double& f1,
                                           double& f2)
{
  double s, c;
  sincos(th, &s, &c);
  if (m.model == kModelCart)
  {  // cart.hpp:94-102: (wheel_radius / 2) * { (u0 + u1) cos, (u0 + u1) sin, (u1 - u0) / wheel_base }
    const double h = __ddiv_rn(m.a, 2.0), su = __dadd_rn(u[0], u[1]);
    f0 = __dmul_rn(h, __dmul_rn(su, c));
    f1 = __dmul_rn(h, __dmul_rn(su, s));
    f2 = __dmul_rn(h, __ddiv_rn(__dsub_rn(u[1], u[0]), m.b));
    return true;
  }
  if (m.model == kModelMecanum)
  {  // omni.hpp:96-107
    const double q = __ddiv_rn(m.a, 4.0);
    const double ss = __dmul_rn(q, s), cc = __dmul_rn(q, c);
    const double l = __ddiv_rn(m.a, __dmul_rn(4.0, __dadd_rn(m.b, m.c)));
    const double spc = __dadd_rn(ss, cc), msc = __dadd_rn(-ss, cc), smc = __dsub_rn(ss, cc);
    f0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(u[0], spc), __dmul_rn(u[1], msc)), __dmul_rn(u[2], spc)), __dmul_rn(u[3], msc));
    f1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(u[0], smc), __dmul_rn(u[1], spc)), __dmul_rn(u[2], smc)), __dmul_rn(u[3], spc));
    f2 = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(-u[0], l), __dmul_rn(u[1], l)), __dmul_rn(u[2], l)), __dmul_rn(u[3], l));
    return true;
  }
  if (m.model == kModelOmni)
  {  // omni.hpp:177-182
    f0 = __dsub_rn(__dmul_rn(u[0], c), __dmul_rn(u[1], s));
    f1 = __dadd_rn(__dmul_rn(u[0], s), __dmul_rn(u[1], c));
    f2 = u[2];
    return true;
  }
  // SimpleCart, cart.hpp:165-173
  f0 = __dmul_rn(u[0], c);
  f1 = __dmul_rn(u[0], s);
  f2 = u[2];
  return fabs(__dsub_rn(u[1], 0.0)) < 1.0e-12;
}

// RungeKutta::solve, forward (integrator.hpp:135-152) with step (:176-184)
__global__ void __launch_bounds__(128) rk4_solve_kernel(const ModelParams m, const int count, const int steps, const double dt,
                                                        const double* __restrict__ x0, const double* __restrict__ ut,
                                                        const long long ut_stride /* 0: one control signal for all */,
                                                        double* __restrict__ xt, int* fault)
{
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= count) return;
  const int nu = model_controls(m.model);
  double x = x0[(size_t)inst * 3 + 0], y = x0[(size_t)inst * 3 + 1], th = x0[(size_t)inst * 3 + 2];
  const double* u_base = ut + (size_t)inst * (size_t)ut_stride;
  double* out = xt + (size_t)inst * (size_t)steps * 3;
  bool ok = true;
  for (int i = 0; i < steps; i++)
  {
    double u[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int c = 0; c < nu; c++) u[c] = u_base[(size_t)i * nu + c];
    double k1[3], k2[3], k3[3], k4[3];
    ok &= model_eval(m, th, u, k1[0], k1[1], k1[2]);
    // x + dt * (0.5 * k)
    ok &= model_eval(m, __dadd_rn(th, __dmul_rn(dt, __dmul_rn(0.5, k1[2]))), u, k2[0], k2[1], k2[2]);
    ok &= model_eval(m, __dadd_rn(th, __dmul_rn(dt, __dmul_rn(0.5, k2[2]))), u, k3[0], k3[1], k3[2]);
    ok &= model_eval(m, __dadd_rn(th, __dmul_rn(dt, k3[2])), u, k4[0], k4[1], k4[2]);
    const double h = __ddiv_rn(dt, 6.0);
    double inc[3];
#pragma unroll
    for (int c = 0; c < 3; c++)  // (dt / 6) * (k1 + 2 k2 + 2 k3 + k4), left to right
      inc[c] = __dmul_rn(h, __dadd_rn(__dadd_rn(__dadd_rn(k1[c], __dmul_rn(2.0, k2[c])), __dmul_rn(2.0, k3[c])), k4[c]));
    x = __dadd_rn(x, inc[0]);
    y = __dadd_rn(y, inc[1]);
    th = normalize_angle_pi(__dadd_rn(th, inc[2]));
    out[(size_t)i * 3 + 0] = x;
    out[(size_t)i * 3 + 1] = y;
    out[(size_t)i * 3 + 2] = th;
  }
  if (!ok) atomicOr(fault, 1);
}

// operator(), fdx, fdu, wheels2Twist of one model for a batch of (x, u): f [count][3], A [count][9] and
// B [count][3 * nu] column-major, vb [count][3]; any output may be null
__global__ void __launch_bounds__(128) model_eval_kernel(const ModelParams m, const int count, const double* __restrict__ xs,
                                                         const double* __restrict__ us, double* __restrict__ f,
                                                         double* __restrict__ A, double* __restrict__ Bm,
                                                         double* __restrict__ vb, int* fault)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int nu = model_controls(m.model);
  const double th = xs[(size_t)i * 3 + 2];
  double u[4] = { 0.0, 0.0, 0.0, 0.0 };
  for (int c = 0; c < nu; c++) u[c] = us[(size_t)i * nu + c];
  double s, c;
  sincos(th, &s, &c);
  if (f)
  {
    double f0, f1, f2;
    if (!model_eval(m, th, u, f0, f1, f2)) atomicOr(fault, 1);
    f[(size_t)i * 3 + 0] = f0;
    f[(size_t)i * 3 + 1] = f1;
    f[(size_t)i * 3 + 2] = f2;
  }
  if (A)
  {
    double* a = A + (size_t)i * 9;
    for (int k = 0; k < 9; k++) a[k] = 0.0;
    double a02, a12;
    if (m.model == kModelCart)
    {  // cart.hpp:114-120
      const double h = __ddiv_rn(m.a, 2.0), su = __dadd_rn(u[0], u[1]);
      a02 = __dmul_rn(__dmul_rn(-h, su), s);
      a12 = __dmul_rn(__dmul_rn(h, su), c);
    }
    else if (m.model == kModelMecanum)
    {  // omni.hpp:120-131
      const double q = __ddiv_rn(m.a, 4.0);
      const double ss = __dmul_rn(q, s), cc = __dmul_rn(q, c);
      const double msc = __dadd_rn(-ss, cc), msmc = __dsub_rn(-ss, cc), spc = __dadd_rn(ss, cc);
      a02 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(u[0], msc), __dmul_rn(u[1], msmc)), __dmul_rn(u[2], msc)), __dmul_rn(u[3], msmc));
      a12 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(u[0], spc), __dmul_rn(u[1], msc)), __dmul_rn(u[2], spc)), __dmul_rn(u[3], msc));
    }
    else if (m.model == kModelOmni)
    {  // omni.hpp:195-196
      a02 = __dsub_rn(__dmul_rn(-u[0], s), __dmul_rn(u[1], c));
      a12 = __dsub_rn(__dmul_rn(u[0], c), __dmul_rn(u[1], s));
    }
    else
    {  // cart.hpp:184-185
      a02 = __dmul_rn(-u[0], s);
      a12 = __dmul_rn(u[0], c);
    }
    a[0 + 3 * 2] = a02;
    a[1 + 3 * 2] = a12;
  }
  if (Bm)
  {
    double* b = Bm + (size_t)i * 3 * nu;
    for (int k = 0; k < 3 * nu; k++) b[k] = 0.0;
    if (m.model == kModelCart)
    {  // cart.hpp:129-142: (wheel_radius / 2) * B
      const double h = __ddiv_rn(m.a, 2.0);
      b[0 + 3 * 0] = __dmul_rn(h, c);
      b[0 + 3 * 1] = __dmul_rn(h, c);
      b[1 + 3 * 0] = __dmul_rn(h, s);
      b[1 + 3 * 1] = __dmul_rn(h, s);
      b[2 + 3 * 0] = __dmul_rn(h, __ddiv_rn(-1.0, m.b));
      b[2 + 3 * 1] = __dmul_rn(h, __ddiv_rn(1.0, m.b));
    }
    else if (m.model == kModelMecanum)
    {  // omni.hpp:139-149
      const double q = __ddiv_rn(m.a, 4.0);
      const double ss = __dmul_rn(q, s), cc = __dmul_rn(q, c);
      const double l = __ddiv_rn(m.a, __dmul_rn(4.0, __dadd_rn(m.b, m.c)));
      const double spc = __dadd_rn(ss, cc), msc = __dadd_rn(-ss, cc), smc = __dsub_rn(ss, cc);
      const double r0[4] = { spc, msc, spc, msc }, r1[4] = { smc, spc, smc, spc }, r2[4] = { -l, l, l, -l };
      for (int k = 0; k < 4; k++)
      {
        b[0 + 3 * k] = r0[k];
        b[1 + 3 * k] = r1[k];
        b[2 + 3 * k] = r2[k];
      }
    }
    else if (m.model == kModelOmni)
    {  // omni.hpp:208-210
      b[0 + 3 * 0] = c;
      b[0 + 3 * 1] = -s;
      b[1 + 3 * 0] = s;
      b[1 + 3 * 1] = c;
      b[2 + 3 * 2] = 1.0;
    }
    else
    {  // cart.hpp:196-202
      b[0 + 3 * 0] = c;
      b[1 + 3 * 0] = s;
      b[2 + 3 * 2] = 1.0;
    }
  }
  if (vb)
  {
    double* v = vb + (size_t)i * 3;
    if (m.model == kModelCart)
    {  // cart.hpp:79-85
      v[0] = __dmul_rn(__ddiv_rn(m.a, 2.0), __dadd_rn(u[0], u[1]));
      v[1] = 0.0;
      v[2] = __dmul_rn(__ddiv_rn(m.a, __dmul_rn(2.0, m.b)), __dsub_rn(u[1], u[0]));
    }
    else if (m.model == kModelMecanum)
    {  // omni.hpp:80-89: (wheel_radius / 4) * Hp * u, Hp = { 1 1 1 1; -1 1 -1 1; -l l l -l }, l = 1 / (bx + by)
      const double l = __ddiv_rn(1.0, __dadd_rn(m.b, m.c)), q = __ddiv_rn(m.a, 4.0);
      const double h0[4] = { q, q, q, q };
      const double h1[4] = { __dmul_rn(q, -1.0), q, __dmul_rn(q, -1.0), q };
      const double h2[4] = { __dmul_rn(q, -l), __dmul_rn(q, l), __dmul_rn(q, l), __dmul_rn(q, -l) };
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int k = 0; k < 4; k++)
      {
        a0 = __dadd_rn(a0, __dmul_rn(h0[k], u[k]));
        a1 = __dadd_rn(a1, __dmul_rn(h1[k], u[k]));
        a2 = __dadd_rn(a2, __dmul_rn(h2[k], u[k]));
      }
      v[0] = a0;
      v[1] = a1;
      v[2] = a2;
    }
    else
    {
      v[0] = u[0];
      v[1] = u[1];
      v[2] = u[2];
    }
  }
}
}  // namespace eb
