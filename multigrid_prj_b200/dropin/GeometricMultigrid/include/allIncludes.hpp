// Drop-in replacement of the reference's GeometricMultigrid/include/allIncludes.hpp.
// Put this directory BEFORE the reference's include directory on the include path: the reference's
// src/main.cpp and src/utilities.cpp then compile unchanged against the B200 library (libmgb200.so).
#ifndef MGB200_DROPIN_ALL_H
#define MGB200_DROPIN_ALL_H

#include <array>
#include <chrono>
#include <cmath>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <tuple>
#include <vector>

#include "utilities.hpp"           // this directory: the reference's declarations + host-sync hooks in the writers
#include "mgb200_gmg_facade.hpp"   // MultiGrid::{SquareDomain, PoissonMatrix, DataVector, smoothers, Residual, Solver, ...}

#endif
