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

#endif
