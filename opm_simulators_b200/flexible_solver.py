"""Host-side mirror of the reference's operator / plugin interface for this path, on top of the
C ABI (include/opmb200.h).  Same names, argument meaning and error behaviour as

  * Opm::PropertyTree                          opm/simulators/linalg/PropertyTree.{hpp,cpp}
  * Dune::MatrixAdapter / GhostLastMatrixAdapter  WellOperators.hpp:400-494
  * Dune::PreconditionerWithUpdate             PreconditionerWithUpdate.hpp:32-41
  * Opm::PreconditionerFactory<Op,Comm>        PreconditionerFactory.hpp:62-159 (create, addCreator)
  * Dune::FlexibleSolver<Operator>             FlexibleSolver.hpp:40-106
  * Opm::AbstractISTLSolver                    AbstractISTLSolver.hpp:43-212 (prepare/solve/iterations)

so that the parity tests read like the reference's own tests.  All numerical work happens in
libopmb200.so on the GPU; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import B200Error, InvalidArgument, MatrixBlockError, SolverAbort  # noqa: F401
from .bcsr import BCSR


# ----------------------------------------------------------------------------------------------
class PropertyTree:
    """Opm::PropertyTree: every leaf is a string (boost::property_tree), keys may be dotted."""

    def __init__(self, source=None):
        if source is None:
            self._t = {}
        elif isinstance(source, PropertyTree):
            self._t = json.loads(json.dumps(source._t))
        elif isinstance(source, dict):
            self._t = json.loads(json.dumps(source))
        elif isinstance(source, str) and source.lstrip().startswith("{"):
            self._t = json.loads(source)
        elif isinstance(source, str):
            with open(source) as f:  # Opm::PropertyTree(const std::string& jsonFile)
                self._t = json.load(f)
        else:
            raise InvalidArgument(_lib.BAD_OPTIONS, f"cannot build a PropertyTree from {type(source)}")

    def _find(self, key):
        node = self._t
        for part in key.split("."):
            if not isinstance(node, dict) or part not in node:
                return None
            node = node[part]
        return node

    def get(self, key, default=None, type=None):
        node = self._find(key)
        if node is None or isinstance(node, dict):
            if default is None:
                raise InvalidArgument(_lib.BAD_OPTIONS, f"PropertyTree: no such key: {key}")
            return default
        conv = type or (default.__class__ if default is not None else str)
        if conv is bool:
            return str(node).lower() in ("true", "1")
        try:
            return conv(node)
        except ValueError as e:
            raise InvalidArgument(_lib.BAD_OPTIONS, f"PropertyTree: cannot convert {key}: {node}") from e

    def put(self, key, value):
        node = self._t
        parts = key.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = value if isinstance(value, str) else repr(value) if isinstance(value, float) else str(value)

    def get_child(self, key):
        node = self._find(key)
        if node is None:
            raise InvalidArgument(_lib.BAD_OPTIONS, f"PropertyTree: no such child: {key}")
        return PropertyTree(node if isinstance(node, dict) else {})

    def get_child_optional(self, key):
        node = self._find(key)
        return None if node is None else PropertyTree(node if isinstance(node, dict) else {})

    def get_child_keys(self):
        return list(self._t.keys())

    def to_json(self) -> str:
        return json.dumps(self._t)


def setup_property_tree(linsolver: str = "ilu0", tol=1e-2, maxiter=200, ilu_relaxation=0.9, verbosity=0):
    """presets of Opm::setupPropertyTree (setupPropertyTree.cpp:417-435 `ilu0`, :491-502 `dilu`);
    a name ending in .json is loaded verbatim (:190-203)"""
    if linsolver.endswith(".json"):
        return PropertyTree(linsolver)
    prm = PropertyTree()
    prm.put("tol", tol)
    prm.put("maxiter", maxiter)
    prm.put("verbosity", verbosity)
    prm.put("solver", "bicgstab")
    if linsolver == "ilu0":
        prm.put("preconditioner.type", "paroverilu0")
        prm.put("preconditioner.relaxation", ilu_relaxation)
        prm.put("preconditioner.ilulevel", 0)
    elif linsolver == "dilu":
        prm.put("preconditioner.type", "dilu")
    else:
        raise InvalidArgument(_lib.BAD_OPTIONS, f"No such linear solver available: {linsolver}")
    return prm


# ----------------------------------------------------------------------------------------------
class Communication:
    """one rank of the NCCL communicator that replaces Dune::OwnerOverlapCopyCommunication's MPI"""

    def __init__(self, rank: int, size: int, unique_id: bytes):
        self.rank, self.size = rank, size
        self._h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        _lib.check(_lib.lib().opmb200_comm_create(rank, size, C.cast(buf, C.c_void_p), C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.lib().opmb200_comm_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    def close(self):
        if self._h:
            _lib.lib().opmb200_comm_destroy(self._h)
            self._h = C.c_void_p()


class MatrixAdapter:
    """Dune::MatrixAdapter<M,X,Y> (serial) / Opm::GhostLastMatrixAdapter (interior_size < n)"""

    def __init__(self, matrix: BCSR, interior_size: int | None = None, comm: Communication | None = None, halo=None):
        self.matrix = matrix
        self.interior_size = matrix.n if interior_size is None else int(interior_size)
        self.comm = comm
        self.halo = halo
        self._solver = None  # set by FlexibleSolver

    def getmat(self) -> BCSR:
        return self.matrix

    def _need(self):
        if self._solver is None:
            raise B200Error(_lib.NOT_PREPARED, "operator is not attached to a FlexibleSolver")
        return self._solver

    def apply(self, x, y):
        """y = A x"""
        _lib.check(_lib.lib().opmb200_op_apply(self._need()._h, _lib.ptr(x), _lib.ptr(y)))

    def applyscaleadd(self, alpha, x, y):
        """y += alpha A x"""
        _lib.check(_lib.lib().opmb200_op_applyscaleadd(self._need()._h, float(alpha), _lib.ptr(x), _lib.ptr(y)))


class WellModelMatrixAdapter(MatrixAdapter):
    """Opm::WellModelMatrixAdapter / WellModelGhostLastMatrixAdapter (WellOperators.hpp:224-287, 300-360): the matrix
    plus the wells as a LinearOperatorExtra.  `wells` = dict(ptr, cells, B, C, Dinv) (see opmb200_set_wells); apply and
    applyscaleadd -- and every operator application inside FlexibleSolver.apply -- compute (A - C^T D^-1 B) x on the
    device.  getmat() is still A, so the preconditioner is A's, as in Flow."""

    def __init__(self, matrix: BCSR, wells: dict, interior_size=None, comm=None, halo=None):
        super().__init__(matrix, interior_size, comm, halo)
        self.wells = wells

    def _attach(self):
        self.set_wells(self.wells)

    def set_wells(self, wells: dict | None):
        """BlackoilWellModel assembles new well equations every Newton iteration"""
        self.wells = wells
        h = self._need()._h
        if not wells or len(wells["ptr"]) <= 1:
            _lib.check(_lib.lib().opmb200_set_wells(h, 0, 0, None, None, None, None, None))
            return
        ptr = np.ascontiguousarray(wells["ptr"], np.int32)
        cells = np.ascontiguousarray(wells["cells"], np.int32)
        Bm, Cm, Di = (np.ascontiguousarray(wells[k], np.float64) for k in ("B", "C", "Dinv"))
        self._keep = (ptr, cells, Bm, Cm, Di)
        _lib.check(_lib.lib().opmb200_set_wells(h, len(ptr) - 1, int(Di.shape[-1]), ptr.ctypes.data, cells.ctypes.data,
                                                Bm.ctypes.data, Cm.ctypes.data, Di.ctypes.data))

    def getNumberOfExtraEquations(self) -> int:
        return 0 if not self.wells else len(self.wells["ptr"]) - 1


class PressureTransferPolicy:
    """The device pieces of Opm::PressureTransferPolicy / Amg::getQuasiImpesWeights around the ILU0/DILU smoother
    (PressureTransferPolicy.hpp:100-162, getQuasiImpesWeights.hpp:64-111; GPU twin gpuistl/detail/
    cpr_amg_operations.cu:35-178).  All vectors natural order, numpy or torch (host or device)."""

    def __init__(self, solver: "FlexibleSolver", pressure_var_index: int = 0, transpose: bool = False, weights=None):
        self.solver, self.p, self.transpose = solver, int(pressure_var_index), bool(transpose)
        self.weights = weights

    def quasi_impes_weights(self, out=None):
        A = self.solver.op.getmat()
        w = np.zeros(A.n * A.b) if out is None else out
        _lib.check(_lib.lib().opmb200_cpr_quasi_impes_weights(self.solver._h, self.p, int(self.transpose), _lib.ptr(w)))
        self.weights = w
        return w

    def calculateCoarseEntries(self, out=None):
        A = self.solver.op.getmat()
        c = np.zeros(A.nnzb) if out is None else out
        _lib.check(_lib.lib().opmb200_cpr_coarse_entries(self.solver._h, _lib.ptr(self.weights), self.p,
                                                         int(self.transpose), _lib.ptr(c)))
        return c

    def moveToCoarseLevel(self, fine, out=None):
        c = np.zeros(self.solver.op.getmat().n) if out is None else out
        _lib.check(_lib.lib().opmb200_cpr_restrict(self.solver._h, _lib.ptr(self.weights), self.p, int(self.transpose),
                                                   _lib.ptr(fine), _lib.ptr(c)))
        return c

    def moveToFineLevel(self, coarse, fine):
        _lib.check(_lib.lib().opmb200_cpr_prolongate(self.solver._h, _lib.ptr(self.weights), self.p, int(self.transpose),
                                                     _lib.ptr(coarse), _lib.ptr(fine)))
        return fine


class PreconditionerWithUpdate:
    """Dune::PreconditionerWithUpdate<X,Y>: pre/apply/post/update/hasPerfectUpdate"""

    type_name = "?"

    def __init__(self, op: MatrixAdapter, prm: PropertyTree):
        self.op, self.prm = op, prm

    def pre(self, x, b):
        pass

    def post(self, x):
        pass

    def apply(self, v, d):
        _lib.check(_lib.lib().opmb200_precond_apply(self.op._need()._h, _lib.ptr(v), _lib.ptr(d)))

    def update(self):
        self.op._need().update()

    def hasPerfectUpdate(self) -> bool:  # DILU.hpp:165, ParallelOverlappingILU0.hpp:147-149
        return True


class PreconditionerFactory:
    """Opm::PreconditionerFactory: a registry type -> creator; unknown types raise
    std::invalid_argument listing the registered ones (PreconditionerFactory_impl.hpp:98-106)."""

    _creators: dict = {}

    @classmethod
    def addCreator(cls, type_name: str, creator):
        cls._creators[type_name.lower()] = creator

    @classmethod
    def create(cls, op: MatrixAdapter, prm: PropertyTree):
        t = prm.get("type", "paroverilu0").lower()
        if t not in cls._creators:
            raise InvalidArgument(_lib.BAD_OPTIONS,
                                  f"Preconditioner type {t} is not registered in the factory. Available types are: "
                                  + " ".join(sorted(cls._creators)))
        return cls._creators[t](op, prm)


def _std_creator(name):
    def make(op, prm):
        p = PreconditionerWithUpdate(op, prm)
        p.type_name = name
        return p
    return make


for _n in ("ilu0", "paroverilu0", "ilun", "dilu", "opmilu0", "opmgpuilu0", "gpuilu0", "gpudilu", "b200ilu0", "b200dilu"):
    PreconditionerFactory.addCreator(_n, _std_creator(_n))


class InverseOperatorResult:
    def __init__(self, r: _lib.Result | None = None):
        self.iterations = r.iterations if r else 0
        self.reduction = r.reduction if r else 0.0
        self.converged = bool(r.converged) if r else False
        self.conv_rate = r.conv_rate if r else 1.0
        self.elapsed = r.elapsed if r else 0.0

    def __repr__(self):
        return (f"InverseOperatorResult(iterations={self.iterations}, reduction={self.reduction:.3e}, "
                f"converged={self.converged}, conv_rate={self.conv_rate:.3f}, elapsed={self.elapsed:.4f})")


class FlexibleSolver:
    """Dune::FlexibleSolver<Operator>(op, [comm,] prm, weightsCalculator, pressureIndex)"""

    def __init__(self, op: MatrixAdapter, prm: PropertyTree | dict | str | None = None, weights_calculator=None,
                 pressure_index: int = 0):
        self.op = op
        self.prm = prm if isinstance(prm, PropertyTree) else PropertyTree(prm)
        child = self.prm.get_child_optional("preconditioner")
        # FlexibleSolver::initOpPrecSp -> PreconditionerFactory::create (unknown type -> invalid_argument)
        self._prec = PreconditionerFactory.create(op, child if child is not None else PropertyTree())
        A = op.getmat()
        self._h = C.c_void_p()
        halo_struct = None
        self._keep = []
        if op.comm is not None:
            h = op.halo
            arrs = [np.ascontiguousarray(h[k], np.int32) for k in
                    ("neighbors", "send_ptr", "send_rows", "recv_ptr", "recv_rows")]
            self._keep = arrs
            ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
            halo_struct = _lib.Halo(len(arrs[0]), ip(arrs[0]), ip(arrs[1]), ip(arrs[2]), ip(arrs[3]), ip(arrs[4]))
        _lib.check(_lib.lib().opmb200_create(self.prm.to_json().encode(), A.b, A.n, A.nnzb, A.rowptr, A.col,
                                             op.interior_size, op.comm._h if op.comm else None,
                                             C.byref(halo_struct) if halo_struct else None, C.byref(self._h)))
        op._solver = self
        self.update()
        if hasattr(op, "_attach"):
            op._attach()

    # ---- Dune::InverseOperator ------------------------------------------------------------------
    def apply(self, x, rhs, reduction: float | None = None) -> InverseOperatorResult:
        """x: in initial guess / out solution; rhs is overwritten with the residual (as Dune does)"""
        res = _lib.Result()
        _lib.check(_lib.lib().opmb200_solve(self._h, _lib.ptr(x), _lib.ptr(rhs),
                                            -1.0 if reduction is None else float(reduction), C.byref(res)))
        return InverseOperatorResult(res)

    def preconditioner(self) -> PreconditionerWithUpdate:
        return self._prec

    def update(self, values=None):
        """GpuSparseMatrixWrapper::updateNonzeroValues + preconditioner().update()"""
        vals = self.op.getmat().val if values is None else values
        _lib.check(_lib.lib().opmb200_update_values(self._h, _lib.ptr(vals)))

    # ---- introspection ------------------------------------------------------------------------------
    def info(self) -> dict:
        i = _lib.Info()
        _lib.check(_lib.lib().opmb200_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def levels(self):
        nl, n = self.info()["n_levels"], self.op.getmat().n
        ptr = np.zeros(nl + 1, np.int32)
        rows = np.zeros(n, np.int32)
        _lib.check(_lib.lib().opmb200_get_levels(self._h, ptr.ctypes.data, rows.ctypes.data))
        return ptr, rows

    def reorder(self):
        n = self.op.getmat().n
        r2n, n2r = np.zeros(n, np.int32), np.zeros(n, np.int32)
        _lib.check(_lib.lib().opmb200_get_reorder(self._h, r2n.ctypes.data, n2r.ctypes.data))
        return r2n, n2r

    def dinv(self):
        A = self.op.getmat()
        d = np.zeros((A.n, A.b, A.b))
        _lib.check(_lib.lib().opmb200_get_dinv(self._h, d.ctypes.data))
        return d

    def ilu0(self):
        A = self.op.getmat()
        lu = np.zeros_like(A.val)
        _lib.check(_lib.lib().opmb200_get_ilu0(self._h, lu.ctypes.data))
        return lu

    def history(self):
        cnt = C.c_int()
        _lib.check(_lib.lib().opmb200_get_history(self._h, None, 0, C.byref(cnt)))
        h = np.zeros(max(cnt.value, 1))
        _lib.check(_lib.lib().opmb200_get_history(self._h, h.ctypes.data, cnt.value, C.byref(cnt)))
        return h[: cnt.value]

    def dot(self, x, y) -> float:
        out = C.c_double()
        _lib.check(_lib.lib().opmb200_dot(self._h, _lib.ptr(x), _lib.ptr(y), C.byref(out)))
        return out.value

    def time_kernel(self, what: int, warmup: int = 3, reps: int = 20):
        ms, nbytes = C.c_double(), C.c_double()
        _lib.check(_lib.lib().opmb200_time_kernel(self._h, what, warmup, reps, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def enable_p2p(self, allgather):
        """switch the collectives from NCCL to the library's own peer-memory kernels.
        `allgather(bytes) -> list[bytes]` gathers one blob per rank in rank order
        (torch.distributed.all_gather_object here, MPI_Allgather in Flow)."""
        blob = C.create_string_buffer(1024)
        _lib.check(_lib.lib().opmb200_p2p_export(self._h, C.cast(blob, C.c_void_p)))
        blobs = b"".join(allgather(blob.raw))
        buf = C.create_string_buffer(blobs, len(blobs))
        _lib.check(_lib.lib().opmb200_p2p_import(self._h, C.cast(buf, C.c_void_p)))

    def timer_start(self):
        _lib.check(_lib.lib().opmb200_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        _lib.check(_lib.lib().opmb200_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            _lib.lib().opmb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NumericalProblem(RuntimeError):
    """Opm::NumericalProblem"""


class ISTLSolverB200:
    """Opm::AbstractISTLSolver back-end (AbstractISTLSolver.hpp:43-212; the reference's GPU twin is
    gpuistl/ISTLSolverGPUISTL.hpp:59-486): prepare(M, b) then solve(x)."""

    def __init__(self, prm=None, relaxed_linear_solver_reduction=1e-2, ignore_convergence_failure=False,
                 comm=None, interior_size=None, halo=None):
        self.prm = prm if isinstance(prm, PropertyTree) else PropertyTree(prm)
        self.relaxed = relaxed_linear_solver_reduction
        self.ignore = ignore_convergence_failure
        self._comm, self._interior, self._halo = comm, interior_size, halo
        self._flex = None
        self._rhs = None
        self._iterations = 0
        self._solve_count = 0
        self.result = None

    def prepare(self, M: BCSR, b):
        """first call: analysis + factorisation; later calls: value refresh + perfect update
        (ISTLSolver.hpp:499-530, gpuistl/ISTLSolverGPUISTL.hpp:425-440)"""
        if self._flex is None:
            self._flex = FlexibleSolver(MatrixAdapter(M, self._interior, self._comm, self._halo), self.prm)
        else:
            self._flex.op.matrix = M
            self._flex.update(M.val)
        self._rhs = b

    def setResidual(self, b):
        self._rhs = b

    def getResidual(self):
        return self._rhs

    def solve(self, x) -> bool:
        self._solve_count += 1
        if int(self.prm.get("verbosity", 0, int)) > 10:
            # ISTLSolver::solve writes the system when verbosity > 10 (ISTLSolver.hpp:433-440 ->
            # WriteSystemMatrixHelper.hpp:63-94, Dune::storeMatrixMarket); scripts/replay_system.py reads it back
            from . import matrixmarket
            import os
            d = self.prm.get("b200.dump_dir", "reports", str)
            os.makedirs(d, exist_ok=True)
            M = self._flex.op.getmat()
            matrixmarket.write_matrix(os.path.join(d, f"prob_{self._solve_count}_matrix_istl.mm"), M)
            matrixmarket.write_vector(os.path.join(d, f"prob_{self._solve_count}_rhs_istl.mm"), np.asarray(self._rhs), M.b)
        self.result = self._flex.apply(x, self._rhs)
        self._iterations = self.result.iterations
        return self.checkConvergence(self.result)

    def iterations(self) -> int:
        return self._iterations

    def getSolveCount(self) -> int:
        return self._solve_count

    def checkConvergence(self, result) -> bool:
        """AbstractISTLSolver::checkConvergence (AbstractISTLSolver.hpp:192-211)"""
        if not result.converged and result.reduction < self.relaxed:
            return True  # "Full linear solver tolerance not achieved" warning in the reference
        if not self.ignore and not result.converged:
            raise NumericalProblem("Convergence failure for linear solver.")
        return result.converged
