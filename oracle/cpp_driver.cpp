// Driver around the UNMODIFIED C++ reference (TEST INFRASTRUCTURE ONLY; built by oracle/Makefile
// into oracle/_ref/, used by oracle/make_golden_cpp.py and tests/test_cpp_reference_pin.py).
//
//   simpleicp_cpp_driver fix.f64 mov.f64 n_fix n_mov correspondences neighbors min_planarity
//                        max_overlap_distance min_change max_iterations out.json
//
// 1. calls the reference's SimpleICP() (c++/src/simpleicp.cpp:8-129) and records the returned H
//    with 17 significant digits (the reference's own screen output, which it prints while
//    running, carries 6 decimals);
// 2. drives the reference's PointCloud / CorrPts classes through the same sequence of calls as
//    simpleicp.cpp:19-80 and records every stage with full precision (selection, normals,
//    matches, distances, kept correspondences, dH, residual statistics).  The final H of this
//    second pass must equal the H of the first bit for bit -- checked here, so the stage dump
//    is known to describe the run SimpleICP() itself makes.
// No algorithm lives in this file: every number comes out of the reference's own functions.
#include "simpleicp.h"
#include "corrpts.h"
#include "pointcloud.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static Eigen::MatrixXd load(const char *path, long n)
{
  std::vector<double> buf(static_cast<size_t>(3 * n));
  FILE *f = std::fopen(path, "rb");
  if (!f || std::fread(buf.data(), sizeof(double), buf.size(), f) != buf.size())
  {
    std::fprintf(stderr, "cannot read %ld points from %s\n", n, path);
    std::exit(2);
  }
  std::fclose(f);
  Eigen::MatrixXd X(n, 3);
  for (long i = 0; i < n; i++)
    for (int j = 0; j < 3; j++)
      X(i, j) = buf[static_cast<size_t>(3 * i + j)];
  return X;
}

static void put_mat4(FILE *o, const char *key, const Eigen::Matrix<double, 4, 4> &H, const char *tail)
{
  std::fprintf(o, "\"%s\": [", key);
  for (int i = 0; i < 4; i++)
    std::fprintf(o, "[%.17g, %.17g, %.17g, %.17g]%s", H(i, 0), H(i, 1), H(i, 2), H(i, 3), i < 3 ? ", " : "");
  std::fprintf(o, "]%s", tail);
}
static void put_ints(FILE *o, const char *key, const std::vector<int> &v, const char *tail)
{
  std::fprintf(o, "\"%s\": [", key);
  for (size_t i = 0; i < v.size(); i++)
    std::fprintf(o, "%d%s", v[i], i + 1 < v.size() ? "," : "");
  std::fprintf(o, "]%s", tail);
}
static void put_vec(FILE *o, const char *key, const Eigen::VectorXd &v, const char *tail)
{
  std::fprintf(o, "\"%s\": [", key);
  for (long i = 0; i < v.size(); i++)
    std::fprintf(o, "%.17g%s", v(i), i + 1 < v.size() ? "," : "");
  std::fprintf(o, "]%s", tail);
}

int main(int argc, char **argv)
{
  if (argc != 12)
  {
    std::fprintf(stderr, "usage: see the header of oracle/cpp_driver.cpp\n");
    return 2;
  }
  const long n_fix = std::atol(argv[3]), n_mov = std::atol(argv[4]);
  const int correspondences = std::atoi(argv[5]), neighbors = std::atoi(argv[6]);
  const double min_planarity = std::atof(argv[7]), max_overlap_distance = std::atof(argv[8]);
  const double min_change = std::atof(argv[9]);
  const int max_iterations = std::atoi(argv[10]);
  const Eigen::MatrixXd X_fix = load(argv[1], n_fix), X_mov = load(argv[2], n_mov);

  // ---- pass 1: the reference's entry point ----
  std::printf("### SimpleICP() begin\n");
  Eigen::Matrix<double, 4, 4> H_api = SimpleICP(X_fix, X_mov, correspondences, neighbors, min_planarity,
                                                max_overlap_distance, min_change, max_iterations);
  std::printf("### SimpleICP() end\n");
  std::fflush(stdout);

  // ---- pass 2: the same calls, stage by stage (simpleicp.cpp:19-80) ----
  FILE *o = std::fopen(argv[11], "w");
  if (!o)
    return 2;
  std::fprintf(o, "{");
  put_mat4(o, "H_api", H_api, ",\n");

  PointCloud pc_fix{X_fix};
  PointCloud pc_mov{X_mov};
  if (max_overlap_distance > 0)
    pc_fix.SelectInRange(pc_mov.X(), max_overlap_distance);
  std::fprintf(o, "\"n_in_range\": %d,\n", static_cast<int>(pc_fix.GetIdxOfSelectedPts().size()));
  pc_fix.SelectNPts(static_cast<uint>(correspondences));
  const std::vector<int> sel = pc_fix.GetIdxOfSelectedPts();
  put_ints(o, "idx_fix", sel, ",\n");
  pc_fix.EstimateNormals(neighbors);
  {
    Eigen::VectorXd nx(sel.size()), ny(sel.size()), nz(sel.size()), pl(sel.size());
    for (size_t i = 0; i < sel.size(); i++)
    {
      nx(static_cast<long>(i)) = pc_fix.nx()(sel[i]);
      ny(static_cast<long>(i)) = pc_fix.ny()(sel[i]);
      nz(static_cast<long>(i)) = pc_fix.nz()(sel[i]);
      pl(static_cast<long>(i)) = pc_fix.planarity()(sel[i]);
    }
    put_vec(o, "nx", nx, ",\n");
    put_vec(o, "ny", ny, ",\n");
    put_vec(o, "nz", nz, ",\n");
    put_vec(o, "planarity", pl, ",\n");
  }

  Eigen::Matrix<double, 4, 4> H_old{Eigen::Matrix<double, 4, 4>::Identity()}, H_new, dH;
  Eigen::VectorXd residual_dists;
  std::vector<double> means, stds;
  bool converged = false;
  std::fprintf(o, "\"iterations\": [\n");
  for (int i = 0; i < max_iterations; i++)
  {
    CorrPts cp = CorrPts(pc_fix, pc_mov);
    cp.Match();
    std::fprintf(o, "%s{", i ? ",\n" : "");
    put_ints(o, "idx_mov_all", cp.idx_pc2(), ", ");
    put_vec(o, "dists_all", cp.dists(), ", ");
    cp.Reject(min_planarity);
    put_ints(o, "idx_fix_kept", cp.idx_pc1(), ", ");
    auto initial_dists{cp.dists()};
    cp.EstimateRigidBodyTransformation(dH, residual_dists);
    pc_mov.Transform(dH);
    H_new = H_old * dH;
    H_old = H_new;
    means.push_back(residual_dists.mean());
    stds.push_back(Std(residual_dists));
    put_mat4(o, "dH", dH, ", ");
    put_mat4(o, "H", H_new, ", ");
    std::fprintf(o, "\"n_kept\": %d, \"initial_mean\": %.17g, \"initial_std\": %.17g, \"mean\": %.17g, \"std\": %.17g}",
                 static_cast<int>(residual_dists.size()), initial_dists.mean(), Std(initial_dists), means.back(),
                 stds.back());
    if (i > 0 && CheckConvergenceCriteria(means, stds, min_change))
    {
      converged = true;
      break;
    }
  }
  std::fprintf(o, "\n],\n\"converged\": %s,\n", converged ? "true" : "false");
  bool same = true;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      same = same && (std::memcmp(&H_new(i, j), &H_api(i, j), sizeof(double)) == 0);
  std::fprintf(o, "\"stage_pass_equals_api\": %s}\n", same ? "true" : "false");
  std::fclose(o);
  if (!same)
  {
    std::fprintf(stderr, "stage pass and SimpleICP() disagree\n");
    return 3;
  }
  return 0;
}
