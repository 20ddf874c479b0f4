// Drop-in counterpart of the reference's AMG/include/CSRMatrix.hpp: the caller-side containers the
// reference's drivers fill (Matrix: one ordered column->value map per row) and read (CSRMatrix).
// They stay on the host -- assembly is the caller's job (AMG/src/main.cpp:34-117) -- and are what the
// facade AMG class converts into the CSR arrays of mgb_amg_create_from_csr().
#ifndef CSR_MATRIX_HPP   // same guard as the reference header: its Utilities.hpp includes its sibling by relative path
#define CSR_MATRIX_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

class Matrix {
    std::vector<std::map<size_t, double>> cells;
    const size_t n_rows, n_cols;
    size_t nnz = 0;

public:
    Matrix(const size_t &rows_, const size_t &cols_) : cells(rows_), n_rows(rows_), n_cols(cols_) {}
    double &at(const size_t &row, const size_t &col)
    {
        if (row >= n_rows || col >= n_cols) throw std::invalid_argument("Index out of range");
        return cells[row][col];
    }
    void count_non_zeros()
    {
        for (const auto &r : cells)
            for (const auto &e : r) if (e.second != 0) ++nnz;
    }
    const size_t non_zeros() { return nnz; }
    const size_t rows() { return n_rows; }
    const size_t cols() { return n_cols; }
    std::vector<std::map<size_t, double>> &data() { return cells; }
    void print()
    {
        for (size_t i = 0; i < n_rows; ++i) { for (size_t j = 0; j < n_cols; ++j) std::cout << at(i, j) << "\t\t"; std::cout << std::endl; }
        std::cout << std::endl;
    }
};

class CSRMatrix {
    size_t n_rows, n_cols;
    std::vector<size_t> row_start;
    std::vector<std::pair<size_t, double>> entries;

public:
    CSRMatrix(const size_t &rows_, const size_t &cols_, const size_t &nnz_) : n_rows(rows_), n_cols(cols_), row_start(rows_ + 1, 0) { entries.reserve(nnz_); }
    explicit CSRMatrix(Matrix &A) : n_rows(A.rows()), n_cols(A.cols()), row_start(A.rows() + 1, 0) { entries.reserve(A.non_zeros()); }
    void copy_from(Matrix &A)
    {
        if (A.rows() != n_rows || A.cols() != n_cols) throw std::invalid_argument("Input matrix doesn't match the size!");
        entries.clear();
        size_t r = 0;
        for (const auto &row : A.data()) {
            row_start[r] = entries.size();
            for (const auto &e : row) if (e.second != 0) entries.push_back(e);     // exact zeros are not stored
            ++r;
        }
        row_start[n_rows] = entries.size();
    }
    const double coeff(const size_t &row, const size_t &col)
    {
        if (row >= n_rows || col >= n_cols) throw std::invalid_argument("Input out of range!!");
        for (size_t k = row_start[row]; k < row_start[row + 1]; ++k) if (entries[k].first == col) return entries[k].second;
        return 0.0;
    }
    const std::vector<std::pair<size_t, double>> nonZerosInRow(const size_t &row)
    {
        if (row >= n_rows) throw std::invalid_argument("Input out of range");
        return std::vector<std::pair<size_t, double>>(entries.begin() + row_start[row], entries.begin() + row_start[row + 1]);
    }
    const size_t rows() { return n_rows; }
    const size_t cols() { return n_cols; }
    void print()
    {
        for (size_t i = 0; i < n_rows; ++i) { for (size_t j = 0; j < n_cols; ++j) std::cout << coeff(i, j) << " "; std::cout << std::endl; }
        std::cout << std::endl;
    }
    std::vector<size_t> component_mask;
    // facade: the raw arrays
    const std::vector<size_t> &row_offsets() const { return row_start; }
    const std::vector<std::pair<size_t, double>> &raw() const { return entries; }
};

#endif
