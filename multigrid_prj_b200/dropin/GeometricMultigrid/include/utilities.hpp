// Drop-in counterpart of the reference's GeometricMultigrid/include/utilities.hpp: same enum, defaults and
// Utils:: entry points (the definitions of Initialization_for_N / init_test_functions still come from the
// reference's own src/utilities.cpp, compiled unchanged).  The only difference: the text writers take their
// argument by const reference and first bring a device-resident vector back to the host, because the
// reference's driver writes `u` straight after the last operator call (src/main.cpp:127-128).
#ifndef MGB200_DROPIN_UTILS_H
#define MGB200_DROPIN_UTILS_H

#include <fstream>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

enum SMOOTHERS { Gauss_Siedel, Jacobi, BiCGSTAB, SMOOTHERS_END };      // utilities.hpp:9-14

#define DEFAULT_N 200
#define DEFAULT_ALPHA 10.0
#define DEFAULT_WIDTH 10.0
#define DEFAULT_LEVEL 2
#define DEFAULT_TEST 1
#define DEFAULT_METHOD Gauss_Siedel

namespace MultiGrid { void sync_to_host(const void *host_vector); }

namespace Utils {

void Initialization_for_N(int argc, char **argv, size_t &N, double &alpha, double &width, int &level,
                          int &functions_to_test, SMOOTHERS &sm);
void init_test_functions(std::function<double(const double, const double)> &f,
                         std::function<double(const double, const double)> &g, int i);

template <class Vector>
void saveVectorOnFile(const Vector &f, std::string fileName)           // utilities.hpp:43-54 (same format)
{
    MultiGrid::sync_to_host(static_cast<const void *>(&f));
    std::ofstream file;
    file.open(fileName, std::ofstream::trunc);
    file << f.size() << std::endl;
    for (size_t i = 0; i < f.size(); i++) file << f[i] << std::endl;
    file.close();
}

template <class SpMat>
void saveMatrixOnFile(SpMat A, std::string fileName)                   // utilities.hpp:28-41 (same format)
{
    std::ofstream file;
    file.open(fileName, std::ofstream::trunc);
    file << A.rows() << " " << A.cols() << " " << A.nonZeros() << std::endl;
    for (size_t i = 0; i < A.rows(); i++)
        for (const auto &j : A.nonZerosInRow(i)) file << i << " " << j << " " << A.coeffRef(i, j) << std::endl;
    file.close();
}

}  // namespace Utils
#endif
