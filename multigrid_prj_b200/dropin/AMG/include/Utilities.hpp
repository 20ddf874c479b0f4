// Drop-in counterpart of the reference's AMG/include/Utilities.hpp.  The problem functions stay the reference's
// (declared here, defined by its own src/Utilities.cpp, compiled unchanged); the smoother class keeps its name,
// constructor and operator* but runs on the device: a CSRMatrix (with or without component_mask) is mirrored as a
// one-level device operator, the iterate stays device-resident across `sol * GS` calls and is copied back when a
// residual is asked for (RestrictionOperator::compute_residual*) or MultiGridAMG::sync_to_host is called.
#ifndef UTILITIES_HPP      // same guard as the reference header (its FEM.hpp includes its sibling by relative path)
#define UTILITIES_HPP

#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

#include "CSRMatrix.hpp"
#include "mgb200.h"

const double boundary_function(const double &x, const double &y);     // Utilities.cpp:3-14
const double forcing_term(const double &x, const double &y);           // Utilities.cpp:16-22
const double alpha(const double &x, const double &y);                   // Utilities.cpp:24-27
int getRandomInit(int max);                                             // Utilities.cpp:30-40
template <typename T> void printVector(std::vector<T> result);

namespace MultiGridAMG {
namespace detail {

inline void ok(int rc) { if (rc != MGB_OK) throw std::runtime_error(std::string("libmgb200: ") + mgb_last_error()); }

// device mirror of one CSRMatrix: x and rhs live in the compact numbering of the matrix rows; a component_mask maps
// row i to entry mask[i] of the caller's (longer) vectors (Utilities.hpp:60-75, AMG.hpp:420-442)
struct DeviceLevel {
    mgb_amg_t h = nullptr;
    size_t n = 0;
    std::vector<size_t> mask;
    const void *rhs_owner = nullptr;
    std::vector<double> *x_owner = nullptr;
    bool x_dirty = false;
    std::vector<double> pack;

    ~DeviceLevel() { if (h) mgb_amg_destroy(h); }
    size_t at(size_t i) const { return mask.empty() ? i : mask[i]; }
    void build(CSRMatrix &A)
    {
        n = A.rows();
        mask = A.component_mask;
        std::vector<int64_t> ptr(n + 1, 0), col;
        std::vector<double> val;
        for (size_t i = 0; i < n; ++i) {
            for (const auto &e : A.nonZerosInRow(i)) { col.push_back((int64_t)e.first); val.push_back(e.second); }
            ptr[i + 1] = (int64_t)col.size();
        }
        std::vector<double> zero(n, 0.0);
        mgb_amg_config c;
        mgb_amg_config_default(&c);
        c.levels = 1;
        ok(mgb_amg_create_from_csr(&c, n, ptr.data(), col.data(), val.data(), zero.data(), &h));
    }
    template <class V> void bind_rhs(V &b)
    {
        if (rhs_owner == &b) return;
        pack.resize(n);
        for (size_t i = 0; i < n; ++i) pack[i] = b[at(i)];
        ok(mgb_amg_set_vector(h, 0, 1, pack.data()));
        rhs_owner = &b;
    }
    void bind_x(std::vector<double> &x)
    {
        if (x_owner == &x) return;
        flush();
        pack.resize(n);
        for (size_t i = 0; i < n; ++i) pack[i] = x.at(at(i));
        ok(mgb_amg_set_vector(h, 0, 0, pack.data()));
        x_owner = &x; x_dirty = false;
    }
    void flush()
    {
        if (!x_owner || !x_dirty) return;
        pack.resize(n);
        ok(mgb_amg_get_vector(h, 0, 0, pack.data()));
        for (size_t i = 0; i < n; ++i) (*x_owner)[at(i)] = pack[i];
        x_dirty = false;
    }
};

inline std::map<CSRMatrix *, std::unique_ptr<DeviceLevel>> &levels()
{
    static std::map<CSRMatrix *, std::unique_ptr<DeviceLevel>> m;
    return m;
}
inline DeviceLevel &level_of(CSRMatrix &A)
{
    auto &p = levels()[&A];
    if (!p || p->n != A.rows() || p->mask != A.component_mask) { p = std::make_unique<DeviceLevel>(); p->build(A); }
    return *p;
}

}  // namespace detail
inline void sync_to_host() { for (auto &kv : detail::levels()) kv.second->flush(); }
}  // namespace MultiGridAMG

template <class Vector>
class SmootherClass {
protected:
    std::function<void(std::vector<double> &)> apply_iteration_to_vec;

public:
    inline friend std::vector<double> &operator*(std::vector<double> &x_k, SmootherClass &B)
    {
        B.apply_iteration_to_vec(x_k);
        return x_k;
    }
};

template <class Vector>
class Gauss_Seidel_iteration : public SmootherClass<Vector> {
    CSRMatrix &m_A;
    Vector &b;

public:
    Gauss_Seidel_iteration(CSRMatrix &A, Vector &f) : m_A(A), b(f)          // Utilities.hpp:81-95
    {
        this->apply_iteration_to_vec = [this](std::vector<double> &sol) {
            auto &L = MultiGridAMG::detail::level_of(m_A);
            L.bind_rhs(b);
            L.bind_x(sol);
            MultiGridAMG::detail::ok(mgb_amg_smooth(L.h, 0, MGB_SMOOTH_GS_LEX, 1));     // exact lexicographic order
            L.x_dirty = true;
        };
    }
    ~Gauss_Seidel_iteration() { MultiGridAMG::sync_to_host(); }
};

#endif
