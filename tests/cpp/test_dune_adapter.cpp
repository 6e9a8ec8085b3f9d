// test_dune_adapter.cpp -- the reference's tests/test_flexiblesolver.cpp:83-130 and
// tests/test_preconditionerfactory.cpp:183-228 replayed through include/opmb200/dune_adapter.hpp
// (Dune-shaped classes -> C ABI -> CUDA).  argv: matrix.mm rhs.mm options.json
#include "stubs/dune_stubs.hpp"

#include "../../include/opmb200/dune_adapter.hpp"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

constexpr int bz = 3;
using Block = Dune::FieldMatrix<double, bz, bz>;
using Matrix = Dune::BCRSMatrix<Block>;
using Vector = Dune::BlockVector<Dune::FieldVector<double, bz>>;
using Operator = Dune::MatrixAdapter<Matrix, Vector, Vector>;

static std::vector<std::string> dataLines(const std::string& path)
{
    std::ifstream f(path);
    if (!f)
        throw std::runtime_error("cannot open " + path);
    std::vector<std::string> out;
    for (std::string l; std::getline(f, l);)
        if (!l.empty() && l[0] != '%')
            out.push_back(l);
    return out;
}

static std::unique_ptr<Matrix> readMatrix(const std::string& path)
{
    auto lines = dataLines(path);
    std::istringstream hs(lines[0]);
    int nr, nc, nnz;
    hs >> nr >> nc >> nnz;
    const int n = nr / bz;
    std::vector<std::map<int, Block>> rows(n);
    for (int k = 1; k <= nnz; ++k) {
        std::istringstream ls(lines[k]);
        int i, j;
        double v;
        ls >> i >> j >> v;
        --i, --j;
        auto& blk = rows[i / bz][j / bz];
        blk[i % bz][j % bz] = v;
    }
    std::vector<int> rp(1, 0), col;
    for (auto& r : rows) {
        for (auto& kv : r)
            col.push_back(kv.first);
        rp.push_back((int)col.size());
    }
    auto A = std::make_unique<Matrix>(rp, col);
    std::size_t k = 0;
    for (auto& r : rows)
        for (auto& kv : r)
            A->blocks()[k++] = kv.second;
    return A;
}

static Vector readVector(const std::string& path)
{
    auto lines = dataLines(path);
    std::istringstream hs(lines[0]);
    int n, one;
    hs >> n >> one;
    Vector v(n / bz);
    for (int i = 0; i < n; ++i)
        v[i / bz][i % bz] = std::stod(lines[1 + i]);
    return v;
}

#define CHECK(cond)                                                                                                    \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__);                                     \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s matrix.mm rhs.mm options.json\n", argv[0]);
        return 2;
    }
    auto A = readMatrix(argv[1]);
    const Vector rhs0 = readVector(argv[2]);
    Opm::PropertyTree prm = Opm::PropertyTree::fromFile(argv[3]);
    Operator op(*A);

    // ---- TestFlexibleSolver (tests/test_flexiblesolver.cpp:112-128), golden values :116-118 ----------
    const double expected[9] = {-1.62493, -1.76435e-06, 1.86991e-10, -458.542, 2.28308e-06, -2.45341e-07,
                                -1.48005, -5.02264e-07, -1.049e-05};
    for (const char* type : {"ilu0", "dilu"}) {
        prm.put("preconditioner.type", std::string(type));
        prm.put("verbosity", 0);
        Opm::b200::Solver<Operator> solver(op, prm.toJson());
        Vector x(rhs0.size()), rhs = rhs0;
        for (std::size_t i = 0; i < x.size(); ++i)
            for (int c = 0; c < bz; ++c)
                x[i][c] = 0.0;
        Dune::InverseOperatorResult res;
        solver.apply(x, rhs, res);
        CHECK(res.converged);
        for (int i = 0; i < 9; ++i) // BOOST_CHECK_CLOSE(sol, expected, 1e-3) [percent]
            CHECK(std::abs(x[i / bz][i % bz] - expected[i]) <= 1e-5 * std::abs(expected[i]));
        // second Newton step: same pattern, new values -> preconditioner().update() (ISTLSolver.hpp:527-528)
        for (auto& b : A->blocks())
            for (int r = 0; r < bz; ++r)
                for (int c = 0; c < bz; ++c)
                    b[r][c] *= 2.0;
        solver.preconditioner().update();
        Vector x2(rhs0.size()), rhs2 = rhs0;
        for (std::size_t i = 0; i < x2.size(); ++i)
            for (int c = 0; c < bz; ++c)
                x2[i][c] = 0.0;
        solver.apply(x2, rhs2, res);
        for (int i = 0; i < 9; ++i)
            CHECK(std::abs(x2[i / bz][i % bz] - 0.5 * expected[i]) <= 1e-5 * std::abs(expected[i]));
        for (auto& b : A->blocks())
            for (int r = 0; r < bz; ++r)
                for (int c = 0; c < bz; ++c)
                    b[r][c] *= 0.5;
        std::printf("b200bicgstab + %s: iterations=%d reduction=%.3e OK\n", type, res.iterations, res.reduction);
    }

    // ---- plugin hook: PreconditionerFactory::addCreator (tests/test_preconditionerfactory.cpp:200-217) -
    using Factory = Opm::PreconditionerFactory<Operator, int>;
    Opm::b200::registerCreators<Factory, Operator>();
    Opm::PropertyTree pp;
    pp.put("type", std::string("b200dilu"));
    auto prec = Factory::create(op, pp);
    CHECK(prec->hasPerfectUpdate());
    Vector v(rhs0.size()), d = rhs0;
    prec->apply(v, d); // block-tridiagonal matrix: DILU is exact, so v solves A v = d
    for (int i = 0; i < 9; ++i)
        CHECK(std::abs(v[i / bz][i % bz] - expected[i]) <= 1e-5 * std::abs(expected[i]));
    prec->update();

    // ---- wells outside the matrix: Solver::setWells, the solution of (A - C^T D^-1 B) x = b satisfies that system ----
    {
        prm.put("preconditioner.type", std::string("dilu"));
        prm.put("tol", 1e-10);
        Opm::b200::Solver<Operator> solver(op, prm.toJson());
        Opm::b200::FlatWells w; // one well, two perforations (cells 0 and 2), dimWells = 2
        w.dimWells = 2;
        w.cells = {0, 2};
        w.ptr = {0, 2};
        for (int p = 0; p < 2; ++p)
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < bz; ++c) {
                    w.B.push_back(1e-3 * (1 + p + r + c));
                    w.C.push_back(2e-3 * (1 + p - r + 2 * c));
                }
        w.Dinv = {0.5, 0.1, -0.2, 0.4};
        solver.setWells(w);
        Vector x(rhs0.size()), rhs = rhs0;
        for (std::size_t i = 0; i < x.size(); ++i)
            for (int c = 0; c < bz; ++c)
                x[i][c] = 0.0;
        Dune::InverseOperatorResult res;
        solver.apply(x, rhs, res);
        CHECK(res.converged);
        // residual of the combined operator, computed here on the host
        Vector y(rhs0.size());
        for (auto row = A->begin(); row != A->end(); ++row) {
            for (int r = 0; r < bz; ++r)
                y[row.index()][r] = 0.0;
            for (auto col = row->begin(); col != row->end(); ++col)
                for (int r = 0; r < bz; ++r)
                    for (int c = 0; c < bz; ++c)
                        y[row.index()][r] += (*col)[r][c] * x[col.index()][c];
        }
        double z1[2] = {0, 0}, z2[2];
        for (int p = 0; p < 2; ++p)
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < bz; ++c)
                    z1[r] += w.B[(p * 2 + r) * bz + c] * x[w.cells[p]][c];
        for (int r = 0; r < 2; ++r)
            z2[r] = w.Dinv[r * 2] * z1[0] + w.Dinv[r * 2 + 1] * z1[1];
        for (int p = 0; p < 2; ++p)
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < bz; ++c)
                    y[w.cells[p]][c] -= w.C[(p * 2 + r) * bz + c] * z2[r];
        double rn = 0, bn = 0;
        for (std::size_t i = 0; i < y.size(); ++i)
            for (int c = 0; c < bz; ++c) {
                rn += (rhs0[i][c] - y[i][c]) * (rhs0[i][c] - y[i][c]);
                bn += rhs0[i][c] * rhs0[i][c];
            }
        CHECK(std::sqrt(rn) <= 1e-8 * std::sqrt(bn));
        solver.clearWells();
        std::printf("b200bicgstab + wells: iterations=%d reduction=%.3e OK\n", res.iterations, res.reduction);
    }

    // ---- error contract: unknown type -> std::invalid_argument (:219-227; PreconditionerFactory_impl.hpp:98-106)
    bool thrown = false;
    try {
        Opm::PropertyTree bad;
        bad.put("preconditioner.type", std::string("not_registered"));
        Opm::b200::Solver<Operator> s(op, bad.toJson());
    } catch (const std::invalid_argument&) {
        thrown = true;
    }
    CHECK(thrown);
    thrown = false;
    try {
        Opm::PropertyTree bad;
        bad.put("solver", std::string("gmres"));
        Opm::b200::Solver<Operator> s(op, bad.toJson());
    } catch (const std::invalid_argument&) {
        thrown = true;
    }
    CHECK(thrown);
    std::printf("dune_adapter: all checks passed\n");
    return 0;
}
