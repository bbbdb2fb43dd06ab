// tests/cpp/adapter_demo.cpp -- exercises the header-only C++ drop-in adapter
// (include/ergodic_exploration_b200/ergodic_control.hpp) the way the
// reference's node mains use ErgodicControl (exploration_omni_node.cpp:139-202).
// Built against the test-only Armadillo stand-in (oracle/shim) because
// Armadillo is not installed in this image; with a real Armadillo the same
// source compiles unchanged.
//
//   adapter_demo cpu            checks that need no GPU; prints OK
//   adapter_demo gpu <0|1>      C1 closed loop on the GPU; prints the numbers
//                               the Python test compares with the oracle
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

#include <ergodic_exploration_b200/ergodic_control.hpp>
#include <ergodic_exploration_b200/collision.hpp>

using namespace ergodic_exploration;

// the caller's own types: a grid with the four getters, and a Collision
namespace ergodic_exploration
{
class Collision
{
};
}  // namespace ergodic_exploration
struct MyGrid
{
  double x0, x1, y0, y1;
  double xmin() const { return x0; }
  double xmax() const { return x1; }
  double ymin() const { return y0; }
  double ymax() const { return y1; }
};

// a map with the reference GridMap's getters (grid.hpp:201-255): a wall and two pillars
struct MyMap
{
  unsigned int nx = 120, ny = 90;
  std::vector<int8_t> cells = std::vector<int8_t>(size_t(120) * 90, 0);
  MyMap()
  {
    for (unsigned int j = 20; j < 100; j++) cells[40 * nx + j] = 100;
    cells[10 * nx + 10] = 100;
    cells[70 * nx + 60] = 77;
    cells[5 * nx + 100] = -1;
  }
  const std::vector<int8_t>& gridData() const { return cells; }
  unsigned int xsize() const { return nx; }
  unsigned int ysize() const { return ny; }
  double resolution() const { return 0.05; }
  double xmin() const { return -1.0; }
  double ymin() const { return 0.5; }
};

#define REQUIRE(cond)                                                     \
  do                                                                      \
  {                                                                       \
    if (!(cond))                                                          \
    {                                                                     \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      std::exit(1);                                                       \
    }                                                                     \
  } while (0)

static void print_vec(const char* name, const double* p, size_t n)
{
  std::printf("%s", name);
  for (size_t i = 0; i < n; i++) std::printf(" %.17g", p[i]);
  std::printf("\n");
}

static int run_cpu()
{
  // test/test_cart.cpp:132-167
  const models::SimpleCart cart;
  const vec x = { 1.0, 2.0, 0.707 }, u = { 0.5, 0.0, 0.01 };
  const vec f = cart(x, u);
  REQUIRE(std::fabs(f(0) - 0.380156) < 1e-6 && std::fabs(f(1) - 0.324777) < 1e-6 && std::fabs(f(2) - 0.01) < 1e-6);
  const mat A = cart.fdx(x, u), B = cart.fdu(x);
  REQUIRE(std::fabs(A(0, 2) + 0.324777) < 1e-6 && std::fabs(A(1, 2) - 0.380156) < 1e-6);
  REQUIRE(std::fabs(B(0, 0) - 0.760313) < 1e-6 && std::fabs(B(1, 0) - 0.649555) < 1e-6 && B(2, 2) == 1.0);
  bool threw = false;
  try
  {
    cart(x, vec({ 0.5, 0.1, 0.0 }));
  }
  catch (const std::invalid_argument&)
  {
    threw = true;
  }
  REQUIRE(threw);
  // test/test_cart.cpp:90-130 and test/test_omni.cpp
  const models::Cart wheels(0.1, 2.0);
  const vec tw = wheels.wheels2Twist(vec({ 0.0, 1.0 }));
  REQUIRE(std::fabs(tw(0) - 0.05) < 1e-9 && tw(1) == 0.0 && std::fabs(tw(2) - 0.025) < 1e-9);
  const models::Mecanum mec(0.1, 0.5, 0.5);
  REQUIRE(mec.fdu(x).n_rows == 3 && mec.fdu(x).n_cols == 4);
  const models::Omni omni;
  REQUIRE(std::fabs(omni(x, vec({ 1.0, 0.0, 0.0 }))(0) - std::cos(0.707)) < 1e-15);
  // Basis tables (basis.cpp:48-77)
  const Basis basis(10.0, 10.0, 4);
  REQUIRE(basis.k()(0, 5) == 1 && basis.k()(1, 5) == 1 && basis.lamdak()(0) == 1.0);
  REQUIRE(std::fabs(basis.lamdak()(5) - 1.0 / std::pow(1.0 + std::sqrt(2.0), 1.5)) < 1e-15);
  // Gaussian (target.hpp:56-107)
  const Gaussian g(vec({ 2.5, 2.5 }), vec({ 1.5, 1.5 }));
  REQUIRE(std::fabs(g(vec({ 2.5, 2.5 })) - 1.0) < 1e-15);
  REQUIRE(std::fabs(g(vec({ 3.5, 2.5 }), vec({ 1.0, 0.0 })) - std::exp(-0.5 * 4.0 / 2.25)) < 1e-15);
  REQUIRE(std::fabs(normalize_angle_PI(3.0 * PI / 2.0) + PI / 2.0) < 1e-12);
  // ctor: one step is rejected exactly like the reference (ergodic_control.hpp:212-216)
  mat Rinv(3, 3, arma::fill::zeros);
  Rinv(0, 0) = 1.0; Rinv(1, 1) = 1.0; Rinv(2, 2) = 2.0;
  const vec umin = { -1.0, -1.0, -2.0 }, umax = { 1.0, 1.0, 2.0 };
  const Collision collision;
  threw = false;
  try
  {
    const ErgodicControl ec(omni, collision, 0.1, 0.1, 0.1, 1.0, 10, 1000, 100, Rinv, umin, umax);
  }
  catch (const std::invalid_argument& e)
  {
    threw = std::strstr(e.what(), "two steps") != nullptr;
  }
  REQUIRE(threw);
  if (eb_device_count() == 0)
  {
    // no GPU: the adapter must fail loudly, never compute on the host
    threw = false;
    try
    {
      const ErgodicControl ec(omni, collision, 0.1, 5.0, 0.1, 1.0, 10, 1000, 100, Rinv, umin, umax);
    }
    catch (const std::runtime_error&)
    {
      threw = true;
    }
    REQUIRE(threw);
    threw = false;
    try
    {
      basis.fourierBasis(vec({ 1.0, 2.0 }));
    }
    catch (const std::runtime_error&)
    {
      threw = true;
    }
    REQUIRE(threw);
  }
  std::printf("OK\n");
  return 0;
}

template <class ModelT>
static int run_gpu(const ModelT& model, const mat& Rinv, const vec& umin, const vec& umax)
{
  const Collision collision;
  // exploration_omni_node.cpp:182-202
  GaussianList gaussians = { Gaussian(vec({ 2.5, 2.5 }), vec({ 1.5, 1.5 })), Gaussian(vec({ 8.5, 2.5 }), vec({ 1.5, 1.5 })) };
  const Target target(gaussians);
  ErgodicControl ec(model, collision, 0.1, 5.0, 0.1, 1.0, 10, 1000000, 100, Rinv, umin, umax);
  ec.setTarget(target);
  const MyGrid grid{ 0.0, 10.0, 0.0, 10.0 };
  vec x = { 5.0, 7.0, 0.3 };
  std::printf("steps %u dt %.17g\n", unsigned(ec.optTraj().n_cols), ec.timeStep());
  for (int step = 0; step < 4; step++)
  {
    ec.addStateMemory(x);  // exploration.hpp:209
    const vec u = ec.control(grid, x);
    print_vec("x", x.memptr(), 3);
    print_vec("u0", u.memptr(), 3);
    const mat ut = ec.controlSignal();
    print_vec("ut", ut.memptr(), ut.n_elem);
    const mat traj = ec.optTraj();
    print_vec("traj", traj.memptr(), traj.n_elem);
    x(0) += 0.1 * u(0);  // any plant will do: the Python side replays the same x
    x(1) += 0.1 * u(1);
    x(2) = normalize_angle_PI(x(2) + 0.1 * u(2));
  }
  {
    // the reference's integrator test (test/test_integrator.cpp:44-73) through the GPU RungeKutta adapter,
    // and a Mecanum rollout for the Python side to compare with the oracle
    const models::Cart cart(0.1, 2.0);
    const mat ut(2, 4, arma::fill::ones);
    const RungeKutta rk4(0.1);
    const mat xt = rk4.solve(cart, vec({ 0.0, 0.0, 0.0 }), ut, 0.4);
    print_vec("rkcart", xt.memptr(), xt.n_elem);
    const models::Mecanum mec(0.05, 0.3, 0.2);
    mat um(4, 6);
    for (unsigned c = 0; c < 6; c++)
      for (unsigned r = 0; r < 4; r++) um(r, c) = 3.0 * std::sin(1.0 + r + 2.0 * c);
    const mat xm = RungeKutta(0.05).solve(mec, vec({ 0.4, -0.2, 1.1 }), um, 0.31);
    print_vec("rkmec", xm.memptr(), xm.n_elem);
  }
  // value semantics: a copy is deep (exploration.hpp:137-138)
  ErgodicControl twin = ec;
  const vec ua = ec.control(grid, x), ub = twin.control(grid, x);
  REQUIRE(ua(0) == ub(0) && ua(1) == ub(1) && ua(2) == ub(2));
  print_vec("phik", ec.targetCoefficients().memptr(), 100);
#ifdef ERGODIC_B200_WITH_ROS
  REQUIRE(ec.path("map").poses.size() == 50 && target.markers("map").markers.size() == 2);
#endif
  // Basis / Target public methods on the GPU
  const Basis basis(10.0, 10.0, 10);
  print_vec("fk", basis.fourierBasis(vec({ 1.25, 7.5 })).memptr(), 100);
  print_vec("dfk", basis.gradFourierBasis(vec({ 1.25, 7.5 })).memptr(), 200);
  mat pts(2, 6);
  for (int i = 0; i < 6; i++)
  {
    pts(0, i) = 1.0 + 1.5 * i;
    pts(1, i) = 9.0 - 1.25 * i;
  }
  print_vec("ck", basis.trajCoeff(pts).memptr(), 100);
  const vec vals = target.fill(vec({ 0.0, 0.0 }), pts);
  print_vec("fill", vals.memptr(), 6);
  print_vec("sc", basis.spatialCoeff(vals, pts).memptr(), 100);
  // batched face
  BatchedErgodicControl bec(model, 3, 0.1, 5.0, 0.1, 1.0, 10, 1000, 100, Rinv, umin, umax);
  bec.setTarget(target);
  mat xb(3, 3);
  for (int i = 0; i < 3; i++)
  {
    xb(0, i) = 2.0 + 3.0 * i;
    xb(1, i) = 8.0 - 2.0 * i;
    xb(2, i) = 0.5 * i;
  }
  vec metric;
  const mat ub3 = bec.control(grid, xb, &metric);
  print_vec("xb", xb.memptr(), 9);
  print_vec("ub", ub3.memptr(), 9);
  print_vec("metric", metric.memptr(), 3);
  // collision checks of the tick that follows control() (exploration.hpp:238)
  {
    const MyMap map;
    const b200::DeviceGrid dgrid(map);
    const b200::Collision col(0.2, 1.0, 0.05, 0.65);
    bool threw = false;
    try
    {
      b200::Collision bad(0.5, 0.2, 0.0, 0.5);
    }
    catch (const std::invalid_argument&)
    {
      threw = true;
    }
    REQUIRE(threw);
    const int B = 24;
    mat p0(3, B), tw(3, B);
    for (int i = 0; i < B; i++)
    {
      p0(0, i) = -0.8 + 0.23 * i;
      p0(1, i) = 0.7 + 0.17 * i;
      p0(2, i) = -3.0 + 0.25 * i;
      tw(0, i) = 0.9 - 0.07 * i;
      tw(1, i) = -0.5 + 0.04 * i;
      tw(2, i) = (i % 4 == 0) ? 0.0 : -1.5 + 0.13 * i;
    }
    const std::vector<int> hit = col.collisionCheck(dgrid, p0);
    const std::vector<int> valid = b200::validate_control(col, dgrid, p0, tw, 0.1, 1.0);
    REQUIRE(col.collisionCheck(dgrid, vec(p0.col(3))) == (hit[3] != 0));
    REQUIRE(b200::validate_control(col, dgrid, vec(p0.col(5)), vec(tw.col(5)), 0.1, 1.0) == (valid[5] != 0));
    print_vec("cpose", p0.memptr(), p0.n_elem);
    print_vec("ctwist", tw.memptr(), tw.n_elem);
    std::vector<double> hd(hit.begin(), hit.end()), vd(valid.begin(), valid.end());
    print_vec("chit", hd.data(), hd.size());
    print_vec("cvalid", vd.data(), vd.size());
    // DynamicWindow (exploration_omni_node.cpp:204-206), one robot and the batch
    const b200::DynamicWindow dwa(col, 0.1, 2.0, 0.2, 2.5, 2.5, 1.0, 1.0, -1.0, 1.0, -1.0, 2.0, -2.0, 3, 8, 5);
    REQUIRE(dwa.steps() == 20);
    mat vref(3, B, arma::fill::zeros);
    const auto [dfound, dtw] = dwa.controlBatch(dgrid, p0, tw, vref);
    const auto [f7, u7] = dwa.control(dgrid, vec(p0.col(7)), vec(tw.col(7)), vec(vref.col(7)));
    REQUIRE(f7 == (dfound[7] != 0) && u7(0) == dtw(0, 7) && u7(1) == dtw(1, 7) && u7(2) == dtw(2, 7));
    const auto [f9, u9] = dwa.control(dgrid, vec(p0.col(9)), vec(tw.col(9)), p0, 0.1);  // any 3 x n path will do
    std::vector<double> fd(dfound.begin(), dfound.end());
    print_vec("dfound", fd.data(), fd.size());
    print_vec("dtwist", dtw.memptr(), dtw.n_elem);
    const double f9d = f9 ? 1.0 : 0.0;
    print_vec("dtraj", &f9d, 1);
    print_vec("dtraju", u9.memptr(), 3);
  }
  std::printf("OK\n");
  return 0;
}

int main(int argc, char** argv)
{
  try
  {
    if (argc < 2 || std::strcmp(argv[1], "cpu") == 0) return run_cpu();
    mat Rinv(3, 3, arma::fill::zeros);
    if (argc > 2 && std::atoi(argv[2]) == 1)
    {
      Rinv(0, 0) = 1.0; Rinv(1, 1) = 1.0; Rinv(2, 2) = 2.0;  // exploration_omni_node.cpp:161-164
      return run_gpu(models::Omni(), Rinv, vec({ -1.0, -1.0, -2.0 }), vec({ 1.0, 1.0, 2.0 }));
    }
    Rinv(0, 0) = 1.0; Rinv(2, 2) = 2.0;  // exploration_cart_node.cpp:156-159
    return run_gpu(models::SimpleCart(), Rinv, vec({ -1.0, 0.0, -2.0 }), vec({ 1.0, 0.0, 2.0 }));
  }
  catch (const std::exception& e)
  {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
}
