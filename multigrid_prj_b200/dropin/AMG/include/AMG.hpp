// Drop-in counterpart of the reference's AMG/include/AMG.hpp: class AMG with the reference's public
// entry points (constructor, apply_AMG, initialization, apply_smoother_operator, compute_residual,
// get_solution, get_x_levels), each a call into the C ABI of include/mgb200.h.  The hierarchy and all
// level vectors live on the device; the reference's stdout lines are kept.
//
// MGB_AMG_MODE=fast selects the multicolour smoother + vector kernels; the default reproduces the
// reference (lexicographic GS, exact order).  MGB_AMG_START=<i> fixes the node the C/F splitting of
// every level starts from (the reference draws it from std::random_device; default n/2).
#ifndef AMG_HPP          // same guard as the reference header
#define AMG_HPP

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "CSRMatrix.hpp"
#include "Utilities.hpp"     // the reference's problem functions (g, f, alpha) -- caller side
#include "mgb200.h"

constexpr double EPSILON = 0.2;      // AMG.hpp:21

class AMG {
    size_t number_of_levels;
    std::vector<int64_t> ptr, col;
    std::vector<double> val, rhs0, x0;
    mgb_amg_t h = nullptr;

    static void ok(int rc) { if (rc != MGB_OK) throw std::runtime_error(std::string("libmgb200: ") + mgb_last_error()); }
    int smoother_kind() const
    {
        const char *m = std::getenv("MGB_AMG_MODE");
        return (m && std::strcmp(m, "fast") == 0) ? MGB_SMOOTH_GS_RB : MGB_SMOOTH_GS_LEX;
    }

public:
    AMG(Matrix &A, std::vector<double> &soll, size_t number_of_levels_, std::vector<double> &rhs_)
        : number_of_levels(number_of_levels_), rhs0(rhs_), x0(soll)          // both copied, as the reference does (AMG.hpp:33-41)
    {
        ptr.assign(A.rows() + 1, 0);
        size_t r = 0;
        for (const auto &row : A.data()) {
            for (const auto &e : row) if (e.second != 0) { col.push_back((int64_t)e.first); val.push_back(e.second); }
            ptr[++r] = (int64_t)col.size();
        }
    }
    ~AMG() { if (h) mgb_amg_destroy(h); }
    AMG(const AMG &) = delete;

    void initialization()                                                       // AMG.cpp:76-120
    {
        if (h) return;
        mgb_amg_config c;
        if (smoother_kind() == MGB_SMOOTH_GS_RB) mgb_amg_config_fast(&c); else mgb_amg_config_default(&c);
        c.levels = (int)number_of_levels;
        if (const char *s = std::getenv("MGB_AMG_START")) for (auto &v : c.start_index) v = std::atoll(s);
        if (const char *d = std::getenv("MGB_DEVICE")) c.device = std::atoi(d);
        ok(mgb_amg_create_from_csr(&c, ptr.size() - 1, ptr.data(), col.data(), val.data(), rhs0.data(), &h));
        ok(mgb_amg_set_vector(h, 0, 0, x0.data()));
        for (size_t l = 1; l < number_of_levels; ++l) {
            size_t n = 0, nf = 0;
            ok(mgb_amg_level_info(h, (int)l, &n, nullptr, nullptr, nullptr, nullptr, nullptr));
            ok(mgb_amg_level_info(h, (int)l - 1, &nf, nullptr, nullptr, nullptr, nullptr, nullptr));
            std::cout << "There are " << n << " coarse nodes at level " << l << std::endl;      // AMG.cpp:85-86
            std::cout << "P size : " << nf << " x " << n << std::endl;                            // AMG.cpp:111
        }
    }
    int apply_smoother_operator(int level, int iter_number)                     // AMG.cpp:236-254
    {
        initialization();
        if (level < 0 || level >= (int)number_of_levels) { std::cerr << "Invalid level: " << level << std::endl; return -1; }
        if (iter_number <= 0) { std::cerr << "Invalid number of iterations: " << iter_number << std::endl; return -1; }
        ok(mgb_amg_smooth(h, level, smoother_kind(), iter_number));
        return 0;
    }
    double compute_residual(int level)                                          // AMG.cpp:256-275
    {
        initialization();
        double norm = 0.;
        ok(mgb_amg_residual(h, level, &norm));
        std::cout << "Residual norm: " << norm << std::endl;
        return norm;
    }
    int apply_AMG()                                                             // AMG.cpp:277-308 (same console trace)
    {
        initialization();
        std::cout << "Initialization done" << std::endl;
        int i;
        for (i = 0; i < (int)number_of_levels - 1; ++i) {
            std::cout << "Applying AMG on level " << i << std::endl << "PRE-SMOOTHING" << std::endl;
            apply_smoother_operator(i, 10);
            std::cout << "COARSENING" << std::endl;
            ok(mgb_amg_restrict(h, i + 1));
        }
        std::cout << "solution on course grid" << std::endl;
        apply_smoother_operator(i, 200);
        std::cout << "PROLUNGATION AND POST-SMOOTHING" << std::endl;
        for (i--; i >= 0; --i) {
            std::cout << "PROLONGATION ON LEVEL " << i << std::endl;
            ok(mgb_amg_prolong(h, i));
            std::cout << "POST-SMOOTHING level: " << i << std::endl;
            apply_smoother_operator(i, 10);
        }
        compute_residual(0);
        std::cout << "AMG applied successfully!" << std::endl;
        return 0;
    }
    std::vector<double> get_x_levels(int level)
    {
        initialization();
        size_t n = 0;
        ok(mgb_amg_level_info(h, level, &n, nullptr, nullptr, nullptr, nullptr, nullptr));
        std::vector<double> x(n);
        ok(mgb_amg_get_vector(h, level, 0, x.data()));
        return x;
    }
    std::vector<double> get_solution() { return get_x_levels(0); }
};

// RestrictionOperator (AMG.hpp:90-452): the setup stages and the residual helpers, as AMG/debugtest.cpp uses them.
class RestrictionOperator {
    static void ok(int rc) { if (rc != MGB_OK) throw std::runtime_error(std::string("libmgb200: ") + mgb_last_error()); }
    struct Handle {                        // host-CSR handle of the C ABI built from a CSRMatrix
        mgb_csr_t h = nullptr;
        explicit Handle(CSRMatrix &A)
        {
            std::vector<int64_t> ptr(A.rows() + 1, 0), col;
            std::vector<double> val;
            for (size_t i = 0; i < A.rows(); ++i) {
                for (const auto &e : A.nonZerosInRow(i)) { col.push_back((int64_t)e.first); val.push_back(e.second); }
                ptr[i + 1] = (int64_t)col.size();
            }
            ok(mgb_csr_create(A.rows(), A.cols(), ptr.data(), col.data(), val.data(), &h));
        }
        ~Handle() { mgb_csr_destroy(h); }
    };
    static std::unique_ptr<CSRMatrix> to_matrix(mgb_csr_t m)
    {
        size_t nr = 0, nc = 0, nnz = 0;
        ok(mgb_csr_info(m, &nr, &nc, &nnz));
        std::vector<int64_t> ptr(nr + 1), col(nnz ? nnz : 1);
        std::vector<double> val(nnz ? nnz : 1);
        ok(mgb_csr_get(m, ptr.data(), col.data(), val.data()));
        Matrix tmp(nr, nc);
        for (size_t i = 0; i < nr; ++i)
            for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) tmp.data()[i][(size_t)col[k]] = val[k];
        tmp.count_non_zeros();
        auto out = std::make_unique<CSRMatrix>(tmp);
        out->copy_from(tmp);
        return out;
    }
    static long start_index()
    {
        const char *s = std::getenv("MGB_AMG_START");
        return s ? std::atol(s) : -1;
    }
    template <bool MASKED>
    double residual(CSRMatrix &A, const std::vector<double> &x, const std::vector<double> &rhs, std::vector<double> &res)
    {
        auto &L = MultiGridAMG::detail::level_of(A);
        L.bind_rhs(rhs);
        L.bind_x(const_cast<std::vector<double> &>(x));
        double norm = 0.;
        ok(mgb_amg_residual(L.h, 0, &norm));
        std::vector<double> r(L.n);
        ok(mgb_amg_get_vector(L.h, 0, 2, r.data()));
        for (size_t i = 0; i < L.n; ++i) res.at(MASKED ? L.at(i) : i) = r[i];
        L.flush();                         // the caller reads x on the host after asking for a residual
        return norm;
    }

public:
    size_t select_coarse_nodes(CSRMatrix &A, std::vector<unsigned char> &coarse_mask)          // AMG.hpp:150-198
    {
        Handle a(A);
        size_t nc = 0;
        coarse_mask.resize(A.rows());
        ok(mgb_amg_select_coarse_nodes(a.h, EPSILON, start_index(), coarse_mask.data(), &nc));
        return nc;
    }
    void build_component_mask(const std::vector<unsigned char> &coarse_mask, std::vector<size_t> &component_mask,
                              const size_t &num_coarse_nodes, std::map<size_t, size_t> &reversed)     // AMG.hpp:201-224
    {
        component_mask.resize(num_coarse_nodes);
        size_t k = 0;
        for (size_t i = 0; i < coarse_mask.size(); ++i)
            if (!(coarse_mask[i] & 0xC0)) { component_mask.at(k) = i; reversed[i] = k; ++k; }
    }
    void build_prolongation_matrix(CSRMatrix &A, std::unique_ptr<CSRMatrix> &P, size_t &, std::vector<unsigned char> &coarse_mask,
                                   std::map<size_t, size_t> &)                                       // AMG.hpp:230-300
    {
        Handle a(A);
        mgb_csr_t p = nullptr;
        ok(mgb_amg_build_prolongation(a.h, EPSILON, coarse_mask.data(), &p));
        P = to_matrix(p);
        mgb_csr_destroy(p);
    }
    void build_coarse_matrix(CSRMatrix &A, CSRMatrix &P, std::unique_ptr<CSRMatrix> &coarse_matrix)  // AMG.hpp:303-369
    {
        Handle a(A), p(P);
        mgb_csr_t c = nullptr;
        ok(mgb_amg_build_coarse_matrix(a.h, p.h, &c));
        coarse_matrix = to_matrix(c);
        mgb_csr_destroy(c);
    }
    void build_coarse_rhs(const std::vector<double> &fine_rhs, CSRMatrix &P, std::vector<double> &coarse_rhs,
                          std::vector<size_t> &component_mask)                                        // AMG.hpp:374-395
    {
        for (size_t j = 0; j < P.rows(); ++j)
            for (const auto &e : P.nonZerosInRow(j)) coarse_rhs.at(component_mask.at(e.first)) += e.second * fine_rhs.at(j);
    }
    double compute_residual(CSRMatrix &A, const std::vector<double> &x, const std::vector<double> &rhs, std::vector<double> &res)
    {
        return residual<false>(A, x, rhs, res);                                                        // AMG.hpp:397-418
    }
    double compute_residual_with_mask(CSRMatrix &A, const std::vector<double> &x, const std::vector<double> &rhs,
                                      std::vector<double> &res)
    {
        return residual<true>(A, x, rhs, res);                                                         // AMG.hpp:420-442
    }
};

#endif
