// Scenarios for tests/test_facade_queue_cpu.py: the reference's operator chains written exactly as its driver writes them
// (GeometricMultigrid/src/main.cpp:41-90), compiled against the facade headers and the recording mock of the C ABI.
#include "allIncludes.hpp"

#include <cstring>
#include <iostream>

using namespace MultiGrid;

int main(int argc, char **argv)
{
    const std::string what = argc > 1 ? argv[1] : "driver";
    const size_t N = 17;
    const int levels = 3;
    std::vector<SquareDomain> domains;
    std::vector<PoissonMatrix<double>> matrici;
    for (int j = 0; j < levels; j++) domains.push_back(SquareDomain(N, 10.0, j));
    for (auto &d : domains) matrici.push_back(PoissonMatrix<double>(d, 1.0));
    std::function<double(double, double)> f = [](double, double) { return 1.0; }, g = [](double, double) { return 0.0; };
    DataVector<double> fvec(domains.front(), f, g);
    std::vector<double> u(matrici.front().rows(), 0.), res(u.size(), 0.), other(u.size(), 0.);
    SawtoothMGIteration<DataVector<double>, Gauss_Seidel_iteration<std::vector<double>>> MG0(matrici, fvec);
    Residual<DataVector<double>> RES(matrici.front(), fvec, res);
    Gauss_Seidel_iteration<DataVector<double>> GS(matrici.front(), fvec);
    std::cout << "== begin " << what << std::endl;
    if (what == "driver") {                       // main.cpp:73-90, three iterations
        u * RES;
        for (int i = 0; i < 3; i++) { u * GS * GS * MG0; u * RES; std::cout << "norm " << RES.Norm() << std::endl; }
    } else if (what == "one_sweep") {             // a chain the queue must NOT fuse: one pre-sweep only
        u * GS * MG0; u * RES;
    } else if (what == "three_sweeps") {
        u * GS * GS * GS * MG0; u * RES;
    } else if (what == "sweeps_then_save") {      // queued sweeps must run before the vector is read back
        u * GS * GS;
        sync_to_host(&u);
        std::cout << "u0 " << u[0] << std::endl;
    } else if (what == "cycle_then_save") {
        u * GS * GS * MG0;
        sync_to_host(&u);
        std::cout << "u0 " << u[0] << std::endl;
    } else if (what == "other_vector") {          // another vector claims the device slot while operators are queued
        u * GS * GS * MG0;
        other * GS;
        sync_to_host(&u); sync_to_host(&other);
        std::cout << "u0 " << u[0] << " other0 " << other[0] << std::endl;
    } else if (what == "bicgstab") {              // the reference's BiCGSTAB class applied to the fine system
        MultiGrid::BiCGSTAB<DataVector<double>> BICG(matrici.front(), fvec, 1e-6);   // (the enum SMOOTHERS has a BiCGSTAB too)
        u * GS * GS;
        u * BICG;
    } else if (what == "residual_of_other") {     // the residual call that follows is about ANOTHER vector: no fusion
        u * GS * GS * MG0;
        other * RES;
        sync_to_host(&u);
        std::cout << "u0 " << u[0] << std::endl;
    }
    std::cout << "== end" << std::endl;
    return 0;
}
