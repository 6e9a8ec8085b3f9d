// dune_adapter.hpp -- the thin C++ -> C-ABI layer that makes libopmb200 a drop-in behind the
// reference's own interfaces.  Header-only; compiles inside an OPM build (needs dune-istl and
// opm-simulators headers) -- in this repository it is compile- and run-checked against the
// minimal stand-ins of tests/cpp/stubs (same class and member names).
//
//   Opm::b200::Solver<Operator>          : Dune::InverseOperator<X,X>
//        what FlexibleSolver::initSolver instantiates for   "solver": "b200bicgstab"
//        (next to "gpubicgstab", FlexibleSolver_impl.hpp:313-321)
//   Opm::b200::Preconditioner<Operator>  : Dune::PreconditionerWithUpdate<X,X>
//        what PreconditionerFactory creates for  "type": "b200dilu" | "b200ilu0"
//        (registered with PreconditionerFactory<Op,Comm>::addCreator, PreconditionerFactory.hpp:116)
//   Opm::b200::registerCreators<Operator, Comm>()   the addCreator calls
//
// Ownership: the Dune matrix and vectors stay with the caller; the handle copies the sparsity once
// (constructor) and the values on every update() (== gpuistl/ISTLSolverGPUISTL.hpp:425-440).
#pragma once

#include "../opmb200.h"

#include <algorithm>
#include <cstdint>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Opm::b200 {

// ---- error conventions ---------------------------------------------------
// OPMB200_* status -> the exception type the reference throws in the same situation.
// MatrixBlockErrorT / SolverAbortT default to the Dune types when dune-istl is present.
template <class MatrixBlockErrorT, class SolverAbortT>
inline void throwOnError(int status)
{
    if (status == OPMB200_SUCCESS)
        return;
    const std::string msg = opmb200_last_error();
    switch (status) {
    case OPMB200_INVALID_ARGUMENT:
    case OPMB200_BAD_OPTIONS:
        throw std::invalid_argument(msg); // PreconditionerFactory_impl.hpp:98-106, FlexibleSolver_impl.hpp:326-329
    case OPMB200_MATRIX_BLOCK_ERROR:
    case OPMB200_DIAGONAL_MISSING:
        throw MatrixBlockErrorT(msg); // rethrown untouched by ISTLSolver::prepare (ISTLSolver.hpp:394-400)
    case OPMB200_SOLVER_ABORT:
        throw SolverAbortT(msg); // derives from Dune::ISTLError -> "Time step too large" chop
    default:
        throw std::runtime_error(msg); // like OPM_GPU_SAFE_CALL
    }
}

// ---- sparsity extraction (== gpuistl GpuSparseMatrix::extractSparsityPattern, GpuSparseMatrix.cpp:113-144)
template <class Matrix>
inline void extractSparsityPattern(const Matrix& A, std::vector<std::int32_t>& rowptr, std::vector<std::int32_t>& colidx)
{
    rowptr.assign(1, 0);
    colidx.clear();
    colidx.reserve(A.nonzeroes());
    for (auto row = A.begin(); row != A.end(); ++row) {
        for (auto col = row->begin(); col != row->end(); ++col)
            colidx.push_back(static_cast<std::int32_t>(col.index()));
        rowptr.push_back(static_cast<std::int32_t>(colidx.size()));
    }
}

// one device handle shared by the operator-facing solver and the preconditioner view
template <class Matrix>
class Handle
{
public:
    using block_type = typename Matrix::block_type;
    static constexpr int blocksize = block_type::rows;

    // serial
    Handle(const Matrix& A, const std::string& jsonOptions)
        : Handle(A, jsonOptions, A.N(), nullptr, nullptr)
    {
    }
    // parallel: interiorSize owner rows first (ISTLSolver.hpp:299-306), NCCL communicator + halo lists
    Handle(const Matrix& A, const std::string& jsonOptions, std::size_t interiorSize, opmb200_comm* comm,
           const opmb200_halo* halo)
        : A_(&A)
    {
        std::vector<std::int32_t> rowptr, colidx;
        extractSparsityPattern(A, rowptr, colidx);
        check(opmb200_create(jsonOptions.empty() ? nullptr : jsonOptions.c_str(), blocksize,
                             static_cast<std::int64_t>(A.N()), static_cast<std::int64_t>(colidx.size()), rowptr.data(),
                             colidx.data(), static_cast<std::int64_t>(interiorSize), comm, halo, &h_));
        parallel_ = comm != nullptr;
        try {
            update(); // may throw Dune::MatrixBlockError (singular pivot block): a recoverable error in Flow
        } catch (...) {
            opmb200_destroy(h_); // no constructor has completed: ~Handle will not run
            h_ = nullptr;
            throw;
        }
    }
    ~Handle() { opmb200_destroy(h_); }
    bool parallel() const { return parallel_; }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;

    // values are contiguous from &A[0][0][0][0] (relied on at gpuistl/GpuSparseMatrix.cpp:164-167)
    void update() { check(opmb200_update_values(h_, &(*A_)[0][0][0][0])); }
    opmb200_solver* get() const { return h_; }
    const Matrix& matrix() const { return *A_; }

    static void check(int status);

private:
    const Matrix* A_;
    opmb200_solver* h_ = nullptr;
    bool parallel_ = false;
};

#ifndef OPMB200_MATRIX_BLOCK_ERROR_T
#define OPMB200_MATRIX_BLOCK_ERROR_T Dune::MatrixBlockError
#endif
#ifndef OPMB200_SOLVER_ABORT_T
#define OPMB200_SOLVER_ABORT_T Dune::SolverAbort
#endif

template <class Matrix>
void Handle<Matrix>::check(int status)
{
    throwOnError<OPMB200_MATRIX_BLOCK_ERROR_T, OPMB200_SOLVER_ABORT_T>(status);
}
// status check for the helpers that have no matrix type at hand
template <>
class Handle<int>
{
public:
    static void checkStatus(int status) { throwOnError<OPMB200_MATRIX_BLOCK_ERROR_T, OPMB200_SOLVER_ABORT_T>(status); }
};

// ---- parallel binding: Dune::OwnerOverlapCopyCommunication -> opmb200_halo + NCCL communicator ---------
// What gpuistl/GpuAwareMPISender.hpp:164-222 derives from comm.remoteIndices(): per neighbour process the
// owner rows this rank sends (entries of the process' send list whose LOCAL attribute is owner) and the copies
// it receives (entries of the receive list whose REMOTE attribute is owner), both in the order of Dune's
// remote index lists (ascending global index on both sides, so the two ends agree).
struct FlatHalo {
    std::size_t interiorSize = 0; // owner rows come first (ISTLSolver.hpp:299-306)
    std::vector<int> neighbors, send_ptr, send_rows, recv_ptr, recv_rows;
    opmb200_halo c {};            // views into the vectors above: valid while *this is alive and not moved from
    void bind()
    {
        c.n_neighbors = static_cast<int>(neighbors.size());
        c.neighbor_rank = neighbors.data();
        c.send_ptr = send_ptr.data();
        c.send_rows = send_rows.data();
        c.recv_ptr = recv_ptr.data();
        c.recv_rows = recv_rows.data();
    }
};

template <class Comm>
inline FlatHalo flattenHalo(const Comm& comm)
{
    constexpr int owner = 1; // Dune::OwnerOverlapCopyAttributeSet::owner
    FlatHalo h;
    std::size_t nOwner = 0, maxOwner = 0;
    bool any = false;
    for (const auto& index : comm.indexSet())
        if (index.local().attribute() == owner) {
            ++nOwner;
            maxOwner = std::max<std::size_t>(maxOwner, index.local().local());
            any = true;
        }
    if (any && maxOwner + 1 != nOwner)
        throw std::invalid_argument("opmb200: owner rows must be numbered first (ghost-last ordering, ISTLSolver.hpp:299-306)");
    h.interiorSize = nOwner;
    h.send_ptr.push_back(0);
    h.recv_ptr.push_back(0);
    for (const auto& process : comm.remoteIndices()) {
        std::vector<int> snd, rcv;
        for (const auto& remote : *process.second.first)
            if (remote.localIndexPair().local().attribute() == owner)
                snd.push_back(static_cast<int>(remote.localIndexPair().local().local()));
        for (const auto& remote : *process.second.second)
            if (remote.attribute() == owner)
                rcv.push_back(static_cast<int>(remote.localIndexPair().local().local()));
        if (snd.empty() && rcv.empty())
            continue;
        h.neighbors.push_back(process.first);
        h.send_rows.insert(h.send_rows.end(), snd.begin(), snd.end());
        h.recv_rows.insert(h.recv_rows.end(), rcv.begin(), rcv.end());
        h.send_ptr.push_back(static_cast<int>(h.send_rows.size()));
        h.recv_ptr.push_back(static_cast<int>(h.recv_rows.size()));
    }
    h.bind();
    return h;
}

// NCCL communicator of `size` ranks: the unique id is made on rank 0 and handed to the others by
// `bcast128(void* buf)` (128 bytes, root 0) -- MPI_Bcast over comm.communicator() in Flow (ncclFromMpi below),
// anything else in a test.  Device binding follows gpuistl/set_device.cpp: rank % number of devices.
template <class Bcast>
inline std::shared_ptr<opmb200_comm> makeComm(int rank, int size, Bcast&& bcast128)
{
    int ndev = 0;
    Handle<int>::checkStatus(opmb200_device_count(&ndev));
    if (ndev > 0)
        Handle<int>::checkStatus(opmb200_set_device(rank % ndev));
    unsigned char id[128] = {0};
    if (rank == 0)
        Handle<int>::checkStatus(opmb200_comm_unique_id(id));
    bcast128(static_cast<void*>(id));
    opmb200_comm* c = nullptr;
    Handle<int>::checkStatus(opmb200_comm_create(rank, size, id, &c));
    return std::shared_ptr<opmb200_comm>(c, [](opmb200_comm* p) { opmb200_comm_destroy(p); });
}

#if defined(HAVE_MPI) && HAVE_MPI
// the one MPI-aware helper: Dune's communicator -> NCCL communicator over the same ranks
template <class Comm>
inline std::shared_ptr<opmb200_comm> ncclFromMpi(const Comm& comm)
{
    MPI_Comm mpi = comm.communicator();
    int rank = 0, size = 1;
    MPI_Comm_rank(mpi, &rank);
    MPI_Comm_size(mpi, &size);
    return makeComm(rank, size, [mpi](void* buf) { MPI_Bcast(buf, 128, MPI_BYTE, 0, mpi); });
}
#endif

// ---- wells kept outside the matrix (matrix-add-well-contributions=false) ----------------------------------
// The flat form of every StandardWell's equations, which is what StandardWellEquations::extract hands the
// reference's GPU bridge (wells/StandardWellEquations.cpp extract(WellContributions&) -> gpubridge/
// WellContributions.cpp addMatrix): per perforation the perforated cell and one dimWells x blocksize block of duneB_
// and duneC_, per well invDuneD_.  Fill it in BlackoilWellModel's loop over the wells after every well assembly and
// hand it to Solver::setWells: the device operator becomes A - sum_w C_w^T D_w^-1 B_w (WellModelMatrixAdapter,
// WellOperators.hpp:224-287) while getmat() -- and with it the preconditioner -- stays A.
struct FlatWells {
    int dimWells = 0;
    std::vector<std::int32_t> ptr {0}, cells;
    std::vector<double> B, C, Dinv;

    // duneB / duneC: one-row BCRS matrices of dimWells x blocksize blocks over the well's perforations (column
    // index = perforation), invDuneD: the dimWells x dimWells inverse block, wellCells[perforation] = local cell
    template <class OffDiagMatrix, class DiagBlock, class Cells>
    void addWell(const OffDiagMatrix& duneB, const OffDiagMatrix& duneC, const DiagBlock& invDuneD, const Cells& wellCells)
    {
        auto rowB = duneB.begin();
        auto rowC = duneC.begin();
        auto colC = rowC->begin();
        for (auto colB = rowB->begin(); colB != rowB->end(); ++colB, ++colC) {
            const auto& blkB = *colB;
            const auto& blkC = *colC;
            if (dimWells == 0)
                dimWells = static_cast<int>(blkB.N());
            cells.push_back(static_cast<std::int32_t>(wellCells[colB.index()]));
            for (std::size_t r = 0; r < blkB.N(); ++r)
                for (std::size_t c = 0; c < blkB.M(); ++c) {
                    B.push_back(blkB[r][c]);
                    C.push_back(blkC[r][c]);
                }
        }
        for (int r = 0; r < dimWells; ++r)
            for (int c = 0; c < dimWells; ++c)
                Dinv.push_back(invDuneD[r][c]);
        ptr.push_back(static_cast<std::int32_t>(cells.size()));
    }
    int numWells() const { return static_cast<int>(ptr.size()) - 1; }
};

// Dune::PreconditionerWithUpdate<X,Y> (PreconditionerWithUpdate.hpp:32-41)
template <class Operator>
class Preconditioner : public Dune::PreconditionerWithUpdate<typename Operator::domain_type, typename Operator::range_type>
{
public:
    using X = typename Operator::domain_type;
    using Y = typename Operator::range_type;
    using Matrix = typename Operator::matrix_type;

    explicit Preconditioner(std::shared_ptr<Handle<Matrix>> h)
        : h_(std::move(h))
    {
    }
    void pre(X&, Y&) override {}
    void post(X&) override {}
    // v = M^-1 d, including the BlockPreconditioner halo copy in parallel
    void apply(X& v, const Y& d) override
    {
        Handle<Matrix>::check(opmb200_precond_apply(h_->get(), &v[0][0], &d[0][0]));
    }
    void update() override { h_->update(); }
    bool hasPerfectUpdate() const override { return true; } // DILU.hpp:165, ParallelOverlappingILU0.hpp:147-149
    // with a communicator this IS the BlockPreconditioner-wrapped preconditioner (ghost restriction and
    // copyOwnerToAll happen inside opmb200_precond_apply): overlapping, like OwningBlockPreconditioner.hpp:81-84
    Dune::SolverCategory::Category category() const override
    {
        return h_->parallel() ? Dune::SolverCategory::overlapping : Dune::SolverCategory::sequential;
    }

private:
    std::shared_ptr<Handle<Matrix>> h_;
};

// Dune::InverseOperator<X,X>: BiCGSTAB + ILU0/DILU entirely on the device
template <class Operator>
class Solver : public Dune::InverseOperator<typename Operator::domain_type, typename Operator::range_type>
{
public:
    using X = typename Operator::domain_type;
    using Matrix = typename Operator::matrix_type;

    // `jsonOptions`: the FlexibleSolver property tree as JSON (prm.write_json); see opmb200_create
    Solver(const Operator& op, const std::string& jsonOptions)
        : h_(std::make_shared<Handle<Matrix>>(op.getmat(), jsonOptions))
        , prec_(std::make_shared<Preconditioner<Operator>>(h_))
    {
    }
    Solver(const Operator& op, const std::string& jsonOptions, std::size_t interiorSize, opmb200_comm* comm,
           const opmb200_halo* halo)
        : h_(std::make_shared<Handle<Matrix>>(op.getmat(), jsonOptions, interiorSize, comm, halo))
        , prec_(std::make_shared<Preconditioner<Operator>>(h_))
    {
    }
    // parallel, from Dune's communication object: FlexibleSolver(op, comm, prm, ...) (FlexibleSolver_impl.hpp:76-92).
    // `nccl` comes from ncclFromMpi(comm) (or makeComm in a test) and is shared by all solvers of the rank.
    template <class Comm>
    Solver(const Operator& op, const Comm& comm, std::shared_ptr<opmb200_comm> nccl, const std::string& jsonOptions)
        : nccl_(std::move(nccl))
    {
        FlatHalo halo = flattenHalo(comm);
        h_ = std::make_shared<Handle<Matrix>>(op.getmat(), jsonOptions, halo.interiorSize, nccl_.get(), &halo.c);
        prec_ = std::make_shared<Preconditioner<Operator>>(h_);
    }

    void apply(X& x, X& b, Dune::InverseOperatorResult& res) override { apply(x, b, -1.0, res); }

    void apply(X& x, X& b, double reduction, Dune::InverseOperatorResult& res) override
    {
        opmb200_result r {};
        const int status = opmb200_solve(h_->get(), &x[0][0], &b[0][0], reduction, &r);
        res.iterations = r.iterations;
        res.reduction = r.reduction;
        res.converged = r.converged != 0;
        res.conv_rate = r.conv_rate;
        res.elapsed = r.elapsed;
        Handle<Matrix>::check(status);
    }

    Dune::SolverCategory::Category category() const override
    {
        return h_->parallel() ? Dune::SolverCategory::overlapping : Dune::SolverCategory::sequential;
    }
    // wells as a LinearOperatorExtra (WellModelMatrixAdapter): every operator application of apply() adds them
    void setWells(const FlatWells& w)
    {
        Handle<Matrix>::check(opmb200_set_wells(h_->get(), w.numWells(), w.dimWells, w.ptr.data(), w.cells.data(),
                                                w.B.data(), w.C.data(), w.Dinv.data()));
    }
    void clearWells() { Handle<Matrix>::check(opmb200_set_wells(h_->get(), 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr)); }
    // FlexibleSolver::preconditioner(): ISTLSolver calls .update() on it every Newton step
    Dune::PreconditionerWithUpdate<X, X>& preconditioner() { return *prec_; }
    std::shared_ptr<Handle<Matrix>> handle() const { return h_; }

private:
    std::shared_ptr<opmb200_comm> nccl_; // keeps the communicator alive as long as the handle
    std::shared_ptr<Handle<Matrix>> h_;
    std::shared_ptr<Preconditioner<Operator>> prec_;
};

// PreconditionerFactory<Operator,Comm>::addCreator(...) for the stand-alone preconditioner use
// (e.g. as CPR fine smoother).  Factory is a template parameter so that this header does not
// depend on opm-simulators' own headers.
template <class Factory, class Operator>
inline void registerCreators()
{
    using Matrix = typename Operator::matrix_type;
    auto make = [](const char* type) {
        return [type](const Operator& op, const auto& prm, const auto& /*weights*/, std::size_t /*pressureIndex*/) {
            std::ostringstream js;
            js.precision(17); // the relaxation factor must survive the round trip through JSON
            js << "{\"preconditioner\": {\"type\": \"" << type << "\", \"relaxation\": \""
               << prm.template get<double>("relaxation", 1.0) << "\"}}";
            auto h = std::make_shared<Handle<Matrix>>(op.getmat(), js.str());
            return std::shared_ptr<Dune::PreconditionerWithUpdate<typename Operator::domain_type,
                                                                  typename Operator::range_type>>(
                std::make_shared<Preconditioner<Operator>>(h));
        };
    };
    Factory::addCreator("b200dilu", make("dilu"));
    Factory::addCreator("b200ilu0", make("ilu0"));
}

// The parallel creators (PreconditionerFactory.hpp:74-77 ParCreator: op, prm, weights, pressureIndex, comm).  The
// preconditioner they return already contains what the reference gets from wrapping a serial preconditioner in
// Dune::BlockPreconditioner (OwningBlockPreconditioner.hpp:31-92): ghost rows restricted, copyOwnerToAll after the
// apply; category() is overlapping.  `nccl` is the rank's communicator (ncclFromMpi(comm)).
template <class Factory, class Operator, class Comm>
inline void registerParallelCreators(std::shared_ptr<opmb200_comm> nccl)
{
    using Matrix = typename Operator::matrix_type;
    auto make = [nccl](const char* type) {
        return [type, nccl](const Operator& op, const auto& prm, const auto& /*weights*/, std::size_t /*pressureIndex*/,
                            const Comm& comm) {
            std::ostringstream js;
            js.precision(17);
            js << "{\"preconditioner\": {\"type\": \"" << type << "\", \"relaxation\": \""
               << prm.template get<double>("relaxation", 1.0) << "\"}}";
            FlatHalo halo = flattenHalo(comm);
            auto h = std::make_shared<Handle<Matrix>>(op.getmat(), js.str(), halo.interiorSize, nccl.get(), &halo.c);
            return std::shared_ptr<Dune::PreconditionerWithUpdate<typename Operator::domain_type,
                                                                  typename Operator::range_type>>(
                std::make_shared<Preconditioner<Operator>>(h));
        };
    };
    Factory::addCreator("b200dilu", make("dilu"));
    Factory::addCreator("b200ilu0", make("ilu0"));
}

} // namespace Opm::b200
