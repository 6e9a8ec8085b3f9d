// dune_stubs.hpp -- minimal stand-ins for the dune-istl / opm-simulators types that
// include/opmb200/dune_adapter.hpp binds to (dune-istl is not installed in this image).  Same
// class and member names as the originals; only what the adapter touches is implemented.
#pragma once
#include <algorithm>
#include <cstddef>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/opmb200/property_tree.hpp"

namespace Dune {

struct ISTLError : std::runtime_error { using std::runtime_error::runtime_error; };
struct MatrixBlockError : ISTLError { using ISTLError::ISTLError; };
struct SolverAbort : ISTLError { using ISTLError::ISTLError; };

struct SolverCategory { enum Category { sequential, nonoverlapping, overlapping }; };

struct InverseOperatorResult {
    int iterations = 0;
    double reduction = 0, conv_rate = 1, elapsed = 0;
    bool converged = false;
    void clear() { *this = InverseOperatorResult(); }
};

template <class K, int n>
struct FieldVector {
    K v[n];
    K& operator[](int i) { return v[i]; }
    const K& operator[](int i) const { return v[i]; }
};

template <class K, int n, int m>
struct FieldMatrix {
    static constexpr int rows = n, cols = m;
    K a[n][m];
    K* operator[](int i) { return a[i]; }
    const K* operator[](int i) const { return a[i]; }
};

template <class B>
class BlockVector
{
public:
    using block_type = B;
    BlockVector() = default;
    explicit BlockVector(std::size_t n) : d_(n) {}
    std::size_t size() const { return d_.size(); }
    B& operator[](std::size_t i) { return d_[i]; }
    const B& operator[](std::size_t i) const { return d_[i]; }
private:
    std::vector<B> d_;
};

// block CSR with contiguous blocks, iterable like Dune::BCRSMatrix
template <class B>
class BCRSMatrix
{
public:
    using block_type = B;
    BCRSMatrix(std::vector<int> rowptr, std::vector<int> col) : rp_(std::move(rowptr)), col_(std::move(col)), val_(col_.size()) {}
    std::size_t N() const { return rp_.size() - 1; }
    std::size_t nonzeroes() const { return col_.size(); }

    struct ColIterator {
        const BCRSMatrix* m; int k;
        std::size_t index() const { return m->col_[k]; }
        const B& operator*() const { return m->val_[k]; }
        ColIterator& operator++() { ++k; return *this; }
        bool operator!=(const ColIterator& o) const { return k != o.k; }
    };
    struct Row {
        const BCRSMatrix* m; int i;
        ColIterator begin() const { return {m, m->rp_[i]}; }
        ColIterator end() const { return {m, m->rp_[i + 1]}; }
        B& operator[](std::size_t j) const
        {
            for (int k = m->rp_[i]; k < m->rp_[i + 1]; ++k)
                if ((std::size_t)m->col_[k] == j)
                    return const_cast<B&>(m->val_[k]);
            throw std::out_of_range("no such block");
        }
    };
    struct RowIterator {
        Row r;
        std::size_t index() const { return r.i; }
        const Row* operator->() const { return &r; }
        const Row& operator*() const { return r; }
        RowIterator& operator++() { ++r.i; return *this; }
        bool operator!=(const RowIterator& o) const { return r.i != o.r.i; }
    };
    RowIterator begin() const { return {{this, 0}}; }
    RowIterator end() const { return {{this, (int)N()}}; }
    Row operator[](std::size_t i) const { return {this, (int)i}; }
    std::vector<B>& blocks() { return val_; }
private:
    std::vector<int> rp_, col_;
    std::vector<B> val_;
};

template <class X, class Y>
struct InverseOperator {
    virtual ~InverseOperator() = default;
    virtual void apply(X& x, Y& b, InverseOperatorResult& res) = 0;
    virtual void apply(X& x, Y& b, double reduction, InverseOperatorResult& res) = 0;
    virtual SolverCategory::Category category() const = 0;
};

template <class X, class Y>
struct Preconditioner {
    virtual ~Preconditioner() = default;
    virtual void pre(X&, Y&) = 0;
    virtual void apply(X& v, const Y& d) = 0;
    virtual void post(X&) = 0;
    virtual SolverCategory::Category category() const = 0;
};

// opm/simulators/linalg/PreconditionerWithUpdate.hpp:32-41
template <class X, class Y>
struct PreconditionerWithUpdate : Preconditioner<X, Y> {
    virtual void update() = 0;
    virtual bool hasPerfectUpdate() const = 0;
};

template <class M, class X, class Y>
class MatrixAdapter
{
public:
    using matrix_type = M;
    using domain_type = X;
    using range_type = Y;
    explicit MatrixAdapter(const M& A) : A_(A) {}
    const M& getmat() const { return A_; }
    SolverCategory::Category category() const { return SolverCategory::sequential; }
private:
    const M& A_;
};

// ---- Dune::OwnerOverlapCopyCommunication<G,L>: only what the halo flattening reads -------------------------
// (dune-istl owneroverlapcopy.hh, dune-common parallel/{indexset,remoteindices,plocalindex}.hh)
struct OwnerOverlapCopyAttributeSet { enum AttributeSet { owner = 1, overlap = 2, copy = 3 }; };

struct ParallelLocalIndexStub {
    std::size_t local_;
    int attribute_;
    std::size_t local() const { return local_; }
    int attribute() const { return attribute_; }
};
struct IndexPairStub {
    int global_;
    ParallelLocalIndexStub local_;
    int global() const { return global_; }
    const ParallelLocalIndexStub& local() const { return local_; }
};
struct RemoteIndexStub {
    int attribute_;          // the attribute of the index on the REMOTE process
    const IndexPairStub* pair_;
    int attribute() const { return attribute_; }
    const IndexPairStub& localIndexPair() const { return *pair_; }
};
using RemoteIndexListStub = std::vector<RemoteIndexStub>;

template <class G, class L>
class OwnerOverlapCopyCommunication
{
public:
    using ParallelIndexSet = std::vector<IndexPairStub>; // sorted by global index, like Dune's
    // rank -> (send list, receive list), as Dune::RemoteIndices: both sorted by global index
    using RemoteIndices = std::map<int, std::pair<RemoteIndexListStub*, RemoteIndexListStub*>>;

    // `globalOf[l]`, `attr[l]` per local index; `ownerRank[l]` the rank that owns it; `peerHas[p]` the global
    // indices present on rank p (what the two sides of Dune's RemoteIndices::rebuild exchange)
    OwnerOverlapCopyCommunication(int rank, const std::vector<int>& globalOf, const std::vector<int>& attr,
                                  const std::map<int, std::vector<std::pair<int, int>>>& peerIndices /* rank -> (global, attr there) */)
        : rank_(rank)
    {
        std::vector<std::size_t> order(globalOf.size());
        for (std::size_t l = 0; l < order.size(); ++l)
            order[l] = l;
        std::sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return globalOf[a] < globalOf[b]; });
        for (std::size_t l : order)
            index_.push_back(IndexPairStub {globalOf[l], ParallelLocalIndexStub {l, attr[l]}});
        std::map<int, const IndexPairStub*> byGlobal;
        for (const auto& ip : index_)
            byGlobal[ip.global()] = &ip;
        for (const auto& peer : peerIndices) {
            auto& lists = store_[peer.first];
            for (const auto& ga : peer.second) { // ascending global index
                auto it = byGlobal.find(ga.first);
                if (it == byGlobal.end())
                    continue; // not shared with that rank
                lists.first.push_back(RemoteIndexStub {ga.second, it->second});
                lists.second.push_back(RemoteIndexStub {ga.second, it->second});
            }
            if (!lists.first.empty())
                remote_[peer.first] = {&lists.first, &lists.second};
        }
    }
    const ParallelIndexSet& indexSet() const { return index_; }
    const RemoteIndices& remoteIndices() const { return remote_; }
    int rank() const { return rank_; }

private:
    int rank_;
    ParallelIndexSet index_;
    std::map<int, std::pair<RemoteIndexListStub, RemoteIndexListStub>> store_;
    RemoteIndices remote_;
};

} // namespace Dune

namespace Opm {
using PropertyTree = opmb200::PropertyTree;

// opm/simulators/linalg/PreconditionerFactory.hpp:62-159 (serial creators only)
template <class Operator, class Comm>
class PreconditionerFactory
{
public:
    using Vector = typename Operator::domain_type;
    using PrecPtr = std::shared_ptr<Dune::PreconditionerWithUpdate<Vector, Vector>>;
    using Creator = std::function<PrecPtr(const Operator&, const PropertyTree&, const std::function<Vector()>&, std::size_t)>;

    static void addCreator(const std::string& type, Creator c) { creators()[type] = std::move(c); }
    static PrecPtr create(const Operator& op, const PropertyTree& prm)
    {
        std::string type = prm.get<std::string>("type", "paroverilu0");
        auto it = creators().find(type);
        if (it == creators().end())
            throw std::invalid_argument("Preconditioner type " + type + " is not registered in the factory.");
        return it->second(op, prm, {}, 0);
    }
private:
    static std::map<std::string, Creator>& creators()
    {
        static std::map<std::string, Creator> m;
        return m;
    }
};
} // namespace Opm
