// stand-in for <nanoflann.hpp>: the KD-tree adaptor API nearest_neighbors.cpp:80-115 calls, answered by an exact brute-force
// search (squared L2 over all columns, ascending distance, equidistant points in ascending index order). A KD-tree returns
// the same neighbours; only the order of exactly equidistant points is implementation-defined (test infrastructure only).
#pragma once
#include <algorithm>
#include <functional>
#include <vector>
namespace nanoflann {
struct SearchParams { SearchParams(int = 32, float = 0, bool = true) {} };
template <class D> class KNNResultSet {
public:
    unsigned long* indices = nullptr;
    D* dists = nullptr;
    size_t capacity;
    explicit KNNResultSet(size_t k) : capacity(k) {}
    void init(unsigned long* i, D* d) { indices = i; dists = d; }
};
template <class M> class KDTreeEigenMatrixAdaptor {
public:
    struct IndexT {
        const M* mat;
        void buildIndex() {}
        template <class RS> void findNeighbors(RS& rs, const typename M::Scalar* q, const SearchParams&) const {
            const int n = mat->rows(), dim = mat->cols();
            std::vector<std::pair<typename M::Scalar, unsigned long>> d(n);
            for (int p = 0; p < n; p++) {
                typename M::Scalar s = 0;
                for (int c = 0; c < dim; c++) { const typename M::Scalar t = q[c] - mat->coeff(p, c); s += t * t; }
                d[p] = {s, (unsigned long)p};
            }
            const size_t k = std::min(rs.capacity, (size_t)n);
            std::partial_sort(d.begin(), d.begin() + k, d.end());
            for (size_t i = 0; i < k; i++) { rs.indices[i] = d[i].second; rs.dists[i] = d[i].first; }
        }
    };
    IndexT impl;
    IndexT* index;
    KDTreeEigenMatrixAdaptor(int /*dim*/, const std::reference_wrapper<const M>& m, int /*leaf*/ = 10) { impl.mat = &m.get(); index = &impl; }
};
}   // namespace nanoflann
