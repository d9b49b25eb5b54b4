// Self-test driver of the Eigen / nanoflann stand-ins (TEST INFRASTRUCTURE ONLY).
// Reads a case description from stdin, runs the stand-in routine, prints the result with 17
// digits; tests/test_cpp_standin.py compares with NumPy / SciPy.  This checks the stand-ins
// themselves, independently of the reference sources compiled against them.
//   eig  n  a11 a12 ... ann          -> eigenvalues ascending, eigenvectors column-major
//   lsq  m n  A (row-major)  b       -> x
//   lin  n lo hi                      -> LinSpaced(n, lo, hi)
//   knn  n k  points (n x 3)  query   -> indices of the k nearest (squared distances after)
#include <Eigen/Dense>
#include <cstdio>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

#include "nanoflann.hpp"

int main()
{
  std::string what;
  while (std::cin >> what)
  {
    if (what == "eig")
    {
      int n;
      std::cin >> n;
      Eigen::MatrixXd A(n, n);
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
          std::cin >> A(i, j);
      Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> es(A);
      for (int i = 0; i < n; i++)
        std::printf("%.17g ", es.eigenvalues()[i]);
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
          std::printf("%.17g ", es.eigenvectors()(i, j));
      std::printf("\n");
    }
    else if (what == "lsq")
    {
      int m, n;
      std::cin >> m >> n;
      Eigen::MatrixXd A(m, n);
      Eigen::VectorXd b(m);
      for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++)
          std::cin >> A(i, j);
      for (int i = 0; i < m; i++)
        std::cin >> b(i);
      Eigen::VectorXd x = A.bdcSvd(Eigen::ComputeThinU | Eigen::ComputeThinV).solve(b);
      for (int j = 0; j < n; j++)
        std::printf("%.17g ", x(j));
      std::printf("\n");
    }
    else if (what == "lin")
    {
      int n;
      double lo, hi;
      std::cin >> n >> lo >> hi;
      Eigen::VectorXd v = Eigen::VectorXd::LinSpaced(n, lo, hi);
      for (int i = 0; i < n; i++)
        std::printf("%.17g ", v(i));
      std::printf("\n");
    }
    else if (what == "knn")
    {
      int n, k;
      std::cin >> n >> k;
      Eigen::MatrixXd X(n, 3);
      for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++)
          std::cin >> X(i, j);
      double q[3];
      std::cin >> q[0] >> q[1] >> q[2];
      typedef nanoflann::KDTreeEigenMatrixAdaptor<Eigen::MatrixXd> kd_tree;
      const Eigen::MatrixXd &Xc = X;
      kd_tree index(3, std::cref(Xc), 10);
      std::vector<size_t> idx(static_cast<size_t>(k));
      std::vector<double> d(static_cast<size_t>(k));
      nanoflann::KNNResultSet<double> rs(static_cast<size_t>(k));
      rs.init(&idx[0], &d[0]);
      index.index_->findNeighbors(rs, q, nanoflann::SearchParameters(10));
      for (int i = 0; i < k; i++)
        std::printf("%zu ", idx[static_cast<size_t>(i)]);
      for (int i = 0; i < k; i++)
        std::printf("%.17g ", d[static_cast<size_t>(i)]);
      std::printf("\n");
    }
  }
  return 0;
}
