// Stand-in for nanoflann.hpp  --  TEST INFRASTRUCTURE ONLY (see Eigen/Dense in this directory).
//
// Implements the part of nanoflann's public interface that /root/reference/c++/src/simpleicp.cpp
// (KnnSearch, lines 131-155) uses: KDTreeEigenMatrixAdaptor<Matrix>(dim, std::cref(mat), leaf),
// its public member index_, index_->findNeighbors(resultSet, query, SearchParameters),
// KNNResultSet<double>(k).init(indices, dists).  Semantics as nanoflann documents them: exact
// k nearest neighbours under the squared Euclidean metric, results sorted by ascending distance.
// Points at EXACTLY equal distance: nanoflann keeps them in tree-traversal order (unspecified);
// this stand-in orders them by ascending index.
#ifndef SICP_ORACLE_NANOFLANN_STANDIN
#define SICP_ORACLE_NANOFLANN_STANDIN

#include <algorithm>
#include <cstddef>
#include <functional>
#include <limits>
#include <numeric>
#include <vector>

namespace nanoflann
{

// nanoflann >= 1.5: SearchParameters(float eps = 0, bool sorted = true).  The reference calls
// SearchParameters(10) (simpleicp.cpp:146) -- written for the older SearchParams(int checks, ...)
// whose first argument was ignored; with nanoflann >= 1.5 that 10 lands in eps and makes the
// library prune with (1 + eps), i.e. an APPROXIMATE search whose misses depend on nanoflann's
// own tree layout.  That is not reproducible without nanoflann itself and is not what the
// algorithm specifies (the Python, MATLAB, Julia and Rust ports all search exactly), so the
// stand-in stores eps and searches exactly.  oracle/linearized_oracle.py does the same.
struct SearchParameters
{
  explicit SearchParameters(float eps_ = 0, bool sorted_ = true) : eps(eps_), sorted(sorted_) {}
  float eps;
  bool sorted;
};

template <typename DistanceType, typename IndexType = size_t, typename CountType = size_t> class KNNResultSet
{
public:
  explicit KNNResultSet(CountType capacity) : indices_(nullptr), dists_(nullptr), capacity_(capacity), count_(0) {}
  void init(IndexType *indices, DistanceType *dists)
  {
    indices_ = indices;
    dists_ = dists;
    count_ = 0;
    if (capacity_)
      dists_[capacity_ - 1] = std::numeric_limits<DistanceType>::max();
  }
  CountType size() const { return count_; }
  bool full() const { return count_ == capacity_; }
  DistanceType worstDist() const { return dists_[capacity_ - 1]; }
  IndexType worstIndex() const { return indices_[capacity_ - 1]; }
  // insertion sort by (distance, index)
  bool addPoint(DistanceType dist, IndexType index)
  {
    CountType i;
    for (i = count_; i > 0; --i)
    {
      if (dists_[i - 1] > dist || (dists_[i - 1] == dist && indices_[i - 1] > index))
      {
        if (i < capacity_)
        {
          dists_[i] = dists_[i - 1];
          indices_[i] = indices_[i - 1];
        }
      }
      else
        break;
    }
    if (i < capacity_)
    {
      dists_[i] = dist;
      indices_[i] = index;
    }
    if (count_ < capacity_)
      count_++;
    return true;
  }

private:
  IndexType *indices_;
  DistanceType *dists_;
  CountType capacity_, count_;
};

template <typename MatrixType> class KDTreeEigenMatrixAdaptor
{
public:
  typedef typename MatrixType::Scalar num_t;

  class index_t
  {
  public:
    index_t(int dim, const MatrixType &m, int leaf) : dim_(dim), m_(m), leaf_(leaf < 1 ? 1 : leaf)
    {
      const size_t n = static_cast<size_t>(m_.rows());
      perm_.resize(n);
      std::iota(perm_.begin(), perm_.end(), size_t(0));
      if (n)
        root_ = build(0, n);
    }

    template <typename RESULTSET>
    bool findNeighbors(RESULTSET &result, const num_t *q, const SearchParameters & /*params*/) const
    {
      // exact search whatever params.eps says (see SearchParameters above)
      if (perm_.empty())
        return false;
      search(root_, q, result);
      return result.full();
    }

  private:
    struct Node
    {
      int axis;         // -1: leaf
      num_t lo, hi;     // split: left subtree has coordinate <= lo, right >= hi
      size_t left, right; // children (inner) or [left, right) range in perm_ (leaf)
    };

    size_t build(size_t b, size_t e)
    {
      Node nd;
      if (e - b <= static_cast<size_t>(leaf_))
      {
        nd.axis = -1;
        nd.lo = nd.hi = 0;
        nd.left = b;
        nd.right = e;
        nodes_.push_back(nd);
        return nodes_.size() - 1;
      }
      int axis = 0;
      num_t best = -1;
      for (int a = 0; a < dim_; a++)
      {
        num_t mn = std::numeric_limits<num_t>::max(), mx = std::numeric_limits<num_t>::lowest();
        for (size_t i = b; i < e; i++)
        {
          const num_t v = m_(static_cast<std::ptrdiff_t>(perm_[i]), a);
          mn = std::min(mn, v);
          mx = std::max(mx, v);
        }
        if (mx - mn > best)
        {
          best = mx - mn;
          axis = a;
        }
      }
      const size_t mid = b + (e - b) / 2;
      std::nth_element(perm_.begin() + static_cast<std::ptrdiff_t>(b), perm_.begin() + static_cast<std::ptrdiff_t>(mid),
                       perm_.begin() + static_cast<std::ptrdiff_t>(e), [&](size_t x, size_t y) {
                         return m_(static_cast<std::ptrdiff_t>(x), axis) < m_(static_cast<std::ptrdiff_t>(y), axis);
                       });
      num_t lo = std::numeric_limits<num_t>::lowest();
      for (size_t i = b; i < mid; i++)
        lo = std::max(lo, m_(static_cast<std::ptrdiff_t>(perm_[i]), axis));
      const num_t hi = m_(static_cast<std::ptrdiff_t>(perm_[mid]), axis);
      const size_t me = nodes_.size();
      nd.axis = axis;
      nd.lo = lo;
      nd.hi = hi;
      nd.left = nd.right = 0;
      nodes_.push_back(nd);
      const size_t l = build(b, mid);
      const size_t r = build(mid, e);
      nodes_[me].left = l;
      nodes_[me].right = r;
      return me;
    }

    template <typename RESULTSET> void search(size_t ni, const num_t *q, RESULTSET &result) const
    {
      const Node &nd = nodes_[ni];
      if (nd.axis < 0)
      {
        for (size_t i = nd.left; i < nd.right; i++)
        {
          const size_t p = perm_[i];
          num_t d = 0;
          for (int a = 0; a < dim_; a++)
          {
            const num_t t = q[a] - m_(static_cast<std::ptrdiff_t>(p), a);
            d += t * t;
          }
          if (d < result.worstDist() || (d == result.worstDist() && (!result.full() || p < result.worstIndex())))
            result.addPoint(d, p);
        }
        return;
      }
      const num_t v = q[nd.axis];
      const num_t dl = v > nd.lo ? v - nd.lo : 0; // distance to the left half-space
      const num_t dr = v < nd.hi ? nd.hi - v : 0;
      const bool left_first = dl <= dr;
      const size_t first = left_first ? nd.left : nd.right, second = left_first ? nd.right : nd.left;
      const num_t d2 = left_first ? dr * dr : dl * dl;
      search(first, q, result);
      if (d2 <= result.worstDist()) // "<=": a tie on the far side may have the lower index
        search(second, q, result);
    }

    int dim_;
    const MatrixType &m_;
    int leaf_;
    std::vector<size_t> perm_;
    std::vector<Node> nodes_;
    size_t root_ = 0;
  };

  KDTreeEigenMatrixAdaptor(const int dimensionality, const std::reference_wrapper<const MatrixType> &mat,
                           const int leaf_max_size = 10)
      : index_(new index_t(dimensionality, mat.get(), leaf_max_size))
  {
  }
  ~KDTreeEigenMatrixAdaptor() { delete index_; }
  KDTreeEigenMatrixAdaptor(const KDTreeEigenMatrixAdaptor &) = delete;
  KDTreeEigenMatrixAdaptor &operator=(const KDTreeEigenMatrixAdaptor &) = delete;

  index_t *index_;
};

} // namespace nanoflann

#endif // SICP_ORACLE_NANOFLANN_STANDIN
