// test_dune_adapter_par.cpp -- the PARALLEL binding of include/opmb200/dune_adapter.hpp against the stand-in
// Dune::OwnerOverlapCopyCommunication of tests/cpp/stubs: flattenHalo (what gpuistl/GpuAwareMPISender.hpp:164-222
// derives from comm.remoteIndices()), makeComm (NCCL id handed round by a caller-supplied broadcast: MPI_Bcast in
// Flow, a file here) and Solver(op, comm, nccl, json) with category() == overlapping.
//   argv: dir rank size flatten|solve
// Reads one rank's ghost-last local system written by tests/test_dune_adapter.py (raw little-endian arrays).
#include "stubs/dune_stubs.hpp"

#include "../../include/opmb200/dune_adapter.hpp"

#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <thread>

constexpr int bz = 3;
using Block = Dune::FieldMatrix<double, bz, bz>;
using Matrix = Dune::BCRSMatrix<Block>;
using Vector = Dune::BlockVector<Dune::FieldVector<double, bz>>;
using Operator = Dune::MatrixAdapter<Matrix, Vector, Vector>;
using Comm = Dune::OwnerOverlapCopyCommunication<int, int>;

template <class T>
static std::vector<T> readRaw(const std::string& path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f)
        throw std::runtime_error("cannot open " + path);
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), bytes);
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s dir rank size flatten|solve\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1], mode = argv[4];
    const int rank = std::atoi(argv[2]), size = std::atoi(argv[3]);
    const std::string pre = dir + "/r" + std::to_string(rank) + "_";
    const auto l2g = readRaw<int>(pre + "l2g.i32"), attr = readRaw<int>(pre + "attr.i32"), peers = readRaw<int>(pre + "peers.i32");
    std::map<int, std::vector<std::pair<int, int>>> peerIndices;
    for (std::size_t k = 1, p = 0; p < (std::size_t)peers[0]; ++p) {
        const int prank = peers[k], cnt = peers[k + 1];
        k += 2;
        for (int i = 0; i < cnt; ++i, k += 2)
            peerIndices[prank].emplace_back(peers[k], peers[k + 1]);
    }
    Comm comm(rank, l2g, attr, peerIndices);
    const Opm::b200::FlatHalo halo = Opm::b200::flattenHalo(comm);

    if (mode == "flatten") { // host only
        std::ofstream o(pre + "halo.txt");
        auto dump = [&](const char* name, const std::vector<int>& v) {
            o << name;
            for (int x : v)
                o << ' ' << x;
            o << '\n';
        };
        o << "interior " << halo.interiorSize << '\n';
        dump("neighbors", halo.neighbors);
        dump("send_ptr", halo.send_ptr);
        dump("send_rows", halo.send_rows);
        dump("recv_ptr", halo.recv_ptr);
        dump("recv_rows", halo.recv_rows);
        return 0;
    }

    // ---- solve: the rank's local system through Solver(op, comm, nccl, json) ----------------------------------
    const auto rowptr = readRaw<int>(pre + "rowptr.i32"), col = readRaw<int>(pre + "col.i32");
    const auto val = readRaw<double>(pre + "val.f64"), rhs = readRaw<double>(pre + "rhs.f64");
    Matrix A(rowptr, col);
    for (std::size_t k = 0; k < col.size(); ++k)
        for (int r = 0; r < bz; ++r)
            for (int c = 0; c < bz; ++c)
                A.blocks()[k][r][c] = val[k * bz * bz + r * bz + c];
    Operator op(A);
    const std::string idfile = dir + "/nccl_id.bin";
    auto nccl = Opm::b200::makeComm(rank, size, [&](void* buf) { // stands in for MPI_Bcast(buf, 128, MPI_BYTE, 0, comm)
        if (rank == 0) {
            std::ofstream(idfile + ".tmp", std::ios::binary).write(static_cast<const char*>(buf), 128);
            std::rename((idfile + ".tmp").c_str(), idfile.c_str());
        } else {
            for (int tries = 0; tries < 3000; ++tries) {
                std::ifstream f(idfile, std::ios::binary);
                if (f && f.read(static_cast<char*>(buf), 128))
                    return;
                std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
            throw std::runtime_error("no NCCL id from rank 0");
        }
    });
    Opm::PropertyTree prm = Opm::PropertyTree::fromFile(dir + "/options.json");
    Opm::b200::Solver<Operator> solver(op, comm, nccl, prm.toJson());
    if (solver.category() != Dune::SolverCategory::overlapping || solver.preconditioner().category() != Dune::SolverCategory::overlapping) {
        std::fprintf(stderr, "category() must be overlapping with a communicator\n");
        return 1;
    }
    Vector x(A.N()), b(A.N());
    for (std::size_t i = 0; i < A.N(); ++i)
        for (int c = 0; c < bz; ++c) {
            x[i][c] = 0.0;
            b[i][c] = rhs[i * bz + c];
        }
    Dune::InverseOperatorResult res;
    solver.apply(x, b, res);
    // second Newton step with the same values: update() + apply must reproduce the first solve
    solver.preconditioner().update();
    Vector x2(A.N()), b2(A.N());
    for (std::size_t i = 0; i < A.N(); ++i)
        for (int c = 0; c < bz; ++c) {
            x2[i][c] = 0.0;
            b2[i][c] = rhs[i * bz + c];
        }
    Dune::InverseOperatorResult res2;
    solver.apply(x2, b2, res2);
    for (std::size_t i = 0; i < A.N(); ++i)
        for (int c = 0; c < bz; ++c)
            if (x2[i][c] != x[i][c]) {
                std::fprintf(stderr, "second solve differs from the first\n");
                return 1;
            }
    std::ofstream o(pre + "x.f64", std::ios::binary);
    for (std::size_t i = 0; i < halo.interiorSize; ++i)
        o.write(reinterpret_cast<const char*>(&x[i][0]), bz * sizeof(double));
    std::printf("rank %d: iterations=%d converged=%d reduction=%.3e\n", rank, res.iterations, (int)res.converged, res.reduction);
    std::ofstream(pre + "result.txt") << res.iterations << ' ' << (int)res.converged << '\n';
    return res.converged ? 0 : 1;
}
