"""
tIGAr's extraction-generator / ExtractedSpline API on the CUDA backend.

Same class and method names, argument meaning and return conventions as the
reference's ``tIGAr/common.py`` (AbstractExtractionGenerator :130-502,
ExtractedSpline :667-1433, AbstractCoordinateChartSpline :1435-1669,
AbstractMultiFieldSpline / EqualOrderSpline / FieldListSpline :1794-1970).
DOLFIN/PETSc objects are replaced by thin device-array wrappers; the numerical
work is done by libtigar_b200.so through tigar_b200.engine.

Two execution modes produce the same IGA system:
  "csr"   reference-faithful: A_FE on the Lagrange mesh, then M^T A M, M^T b
  "fused" element-fused sum_e M_e^T K_e M_e straight into the IGA matrix
          (never forms A_FE or M; the only option at 256^3+).
"""
import abc
import math
import os
import sys
import weakref

import numpy as np

from . import dev
from . import symbolic as S
from . import ufl_lite as U
from .bsplines import (AbstractScalarBasis, AbstractControlMesh, BSpline, DofList, DOLFIN_EPS,
                       USE_RECT_ELEM_DEFAULT, near)
from .engine import TensorPatch, WinMatrix, IGNORE_EPS
from ._lib import lib, check

INDEX_TYPE = 'int32'
DEFAULT_PREALLOC = 500
DEFAULT_BASIS_FUNC_IGNORE_EPS = IGNORE_EPS
EXTRACTION_DATA_FILE = "extraction-data.npz"      # reference: .h5 (no HDF5 here)
EXTRACTION_INFO_FILE = "extraction-info.txt"
EXTRACTION_ZERO_DOFS_FILE = "zero-dofs.dat"
EXTRACTION_MAT_FILE = "extraction-mat.dat"
EXTRACTION_MAT_FILE_CTRL = "extraction-mat-ctrl.dat"
USE_DG_DEFAULT = True
FORM_MT = False


class _Comm(object):
    """Stand-in for an MPI communicator: the torch.distributed world (one
    process per GPU) or a single process."""

    def __init__(self, world=True):
        self.world = world

    def _dist(self):
        import torch.distributed as dist
        return dist if (self.world and dist.is_available() and dist.is_initialized()) else None

    @property
    def rank(self):
        d = self._dist()
        return d.get_rank() if d else 0

    @property
    def size(self):
        d = self._dist()
        return d.get_world_size() if d else 1


worldcomm = _Comm(True)
selfcomm = _Comm(False)


class _MPI(object):
    comm_world = worldcomm
    comm_self = selfcomm

    @staticmethod
    def rank(comm):
        return comm.rank

    @staticmethod
    def size(comm):
        return comm.size

    @staticmethod
    def barrier(comm):
        d = comm._dist()
        if d:
            d.barrier()


MPI = _MPI()
mpisize = MPI.size(worldcomm)
mpirank = MPI.rank(worldcomm)
DEFAULT_DO_PERMUTATION = mpisize > 8


class SubDomain(object):
    def inside(self, x, on_boundary):
        return False


# ---------------------------------------------------------------- vectors
class DeviceVector(object):
    """Minimal stand-in for a DOLFIN PETScVector living in HBM.  ``distributed=True`` marks a
    slab of a row-distributed vector: its norm is reduced over the ranks (a PETSc Vec's norm
    is global), so every rank of a Newton loop takes the same branch."""

    def __init__(self, t, distributed=False, owner=None):
        self.t = t
        self.distributed = distributed
        self.owner = owner            # the Function whose FE coefficients this vector is

    def get_local(self):
        a = dev.to_np(self.t)
        # an independent host array (to_np of a large vector already is a fresh pinned block)
        return a if self.t.numel() * 8 >= (1 << 20) and self.t.is_cuda else a.copy()

    def set_local(self, a):
        self.t.copy_(dev.from_np(np.asarray(a, dtype=np.float64)))
        if self.owner is not None:    # u.vector().set_local(...): the FE data is now the truth
            self.owner._fe_written()

    def size(self):
        return self.t.numel()

    def __len__(self):
        return self.t.numel()

    def __getitem__(self, i):
        return self.get_local()[i]

    def __setitem__(self, i, v):
        if isinstance(v, DeviceVector):
            v = v.get_local()
        a = self.get_local()
        a[i] = v
        self.set_local(a)

    def norm(self, kind="l2"):
        out = dev.zeros(1)
        scratch = dev.empty(lib.tg_cg_scratch_len())
        check(lib.tg_dot(dev.ptr(self.t), dev.ptr(self.t), self.t.numel(), dev.ptr(scratch),
                         dev.ptr(out), dev.stream()))
        if self.distributed:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(out)
        return math.sqrt(float(out.item()))

    def vec(self):
        return self


def norm(v, kind="l2"):
    return v.norm(kind)


class FunctionSpace(object):
    """FE function space on the extraction mesh (CG Q_pf, one cell per knot
    span); ``nfields`` > 1 stands for the MixedElement of common.py:337-351."""

    def __init__(self, owner, nfields=1, control=False):
        self.owner = owner
        self.nfields = nfields
        self.control = control

    def dim(self):
        return self.owner.patch().n_fe * self.nfields

    def mesh(self):
        return self.owner.mesh

    def num_sub_spaces(self):
        return 0 if self.nfields == 1 else self.nfields

    def sub(self, i):
        """Scalar space of field ``i`` (all fields share the control mesh's spline)."""
        if not 0 <= i < self.nfields:
            raise IndexError("field %d of a %d-field space" % (i, self.nfields))
        if self.nfields == 1:
            return self
        if getattr(self, "_subs", None) is None:
            self._subs = [FunctionSpace(self.owner, 1) for _ in range(self.nfields)]
        return self._subs[i]


_fid_counter = [0]
_functions = weakref.WeakValueDictionary()


class Function(U.Tensor):
    """FE function on the extraction mesh.  Holds FE nodal coefficients
    (``fe``) and, when it was produced from IGA DoFs, those too (``iga``)."""

    def __new__(cls, V):
        if cls is Function and V.nfields != 1:
            return object.__new__(VectorFunction)
        return object.__new__(cls)

    def __init__(self, V):
        _fid_counter[0] += 1
        self.fid = _fid_counter[0]
        self.V = V
        self._fe = None
        self.iga = None
        U.Tensor.__init__(self, U.Scalar.coef(S.jet(self.fid, 0, (0, 0, 0))))
        _functions[self.fid] = self

    # data access --------------------------------------------------------
    def fe_tensor(self):
        if self._fe is None:
            if self.iga is not None:
                self._fe = self.V.owner.M_matrix().matvec(self.iga)
            else:
                self._fe = dev.zeros(self.V.owner.patch().n_fe)
        return self._fe

    def set_iga(self, t):
        self.iga = t
        self._fe = None

    def vector(self):
        return DeviceVector(self.fe_tensor(), owner=self)

    def _fe_written(self):
        """The FE coefficients were written through ``u.vector()`` (the reference idiom): the
        IGA DoFs are stale and are re-derived on demand (FEtoIGA, common.py:968-993) instead of
        being read silently by the element-fused / matrix-free paths."""
        self.iga = None

    def assign(self, other):
        """``u.assign(v)`` or ``u.assign(c1*v1 + c2*v2 + ...)`` (dolfin accepts linear
        combinations of Functions of the same space; timeIntegration.py uses them)."""
        if isinstance(other, Function):
            self.iga = None if other.iga is None else other.iga.clone()
            self._fe = None if other._fe is None else other._fe.clone()
            return
        combo = linear_combination(other)
        funcs = [(c, _functions[fid]) for fid, c in combo]
        if all(f.iga is not None for _, f in funcs):
            acc = None
            for c, f in funcs:
                acc = f.iga * c if acc is None else acc + f.iga * c
            self.set_iga(acc)
            return
        acc = None
        for c, f in funcs:
            acc = f.fe_tensor() * c if acc is None else acc + f.fe_tensor() * c
        self.iga = None
        self._fe = acc

    def rename(self, *a):
        pass

    def function_space(self):
        return self.V


class VectorFunction(Function):
    """Function on a multi-field space (``Function(V)`` with ``V.nfields > 1``): a
    vector of scalar component Functions, one per field, all on the control mesh's
    spline.  IGA DoFs are field-major (``globalDof``, common.py:254-262)."""

    def __init__(self, V):
        self.V = V
        self.fid = None
        self.comps = [Function(V.sub(i)) for i in range(V.nfields)]
        arr = np.empty((V.nfields,), dtype=object)
        for i, c in enumerate(self.comps):
            arr[i] = c.a[()]
        U.Tensor.__init__(self, arr)

    def sub(self, i):
        return self.comps[i]

    def split(self):
        return tuple(self.comps)

    def fid_fields(self):
        return {c.fid: i for i, c in enumerate(self.comps)}

    @property
    def iga(self):
        import torch
        if any(c.iga is None for c in self.comps):
            return None
        return torch.cat([c.iga for c in self.comps])

    @property
    def _fe(self):
        import torch
        return torch.cat([c.fe_tensor() for c in self.comps])

    def fe_tensor(self):
        return self._fe

    def set_iga(self, t):
        n = t.numel() // len(self.comps)
        if n * len(self.comps) != t.numel():
            raise ValueError("IGA vector length is not a multiple of the field count")
        for i, c in enumerate(self.comps):
            c.set_iga(t[i * n:(i + 1) * n])

    def assign(self, other):
        other = U.as_tensor(other)
        if other.a.shape != (len(self.comps),):
            raise ValueError("assign: expected a %d-vector" % len(self.comps))
        for i, c in enumerate(self.comps):
            o = other[i]
            src = None
            if isinstance(other, VectorFunction):
                src = other.comps[i]
            c.assign(src if src is not None else o)


def split(u):
    """dolfin ``split``: the field components of a multi-field Function / argument."""
    if isinstance(u, VectorFunction):
        return u.split()
    return tuple(u[i] for i in range(len(u)))


def linear_combination(expr):
    """[(fid, coefficient)] of an expression that is a linear combination, with
    constant coefficients, of (undifferentiated) Functions; ValueError otherwise."""
    sc = U.as_tensor(expr)
    if sc.a.ndim != 0:
        raise ValueError("assign: scalar expressions only")
    sc = sc.a[()]
    if not sc.is_coef():
        raise ValueError("assign: the expression contains trial/test functions")
    n = S.freeze_params(sc.node())
    leaves = S.jets_of([n])
    out = []
    for leaf in leaves:
        fid, comp, al = leaf.args
        if comp != 0 or any(al):
            raise ValueError("assign: derivatives of Functions are not a nodal combination")
        c = S.diff_leaf(n, leaf)
        if not c.is_const():
            raise ValueError("assign: the expression is not linear in its Functions")
        if fid not in _functions:
            raise ValueError("assign: unknown Function in the expression")
        out.append((fid, float(c.args[0])))
    rest = S.substitute(n, {leaf: S.ZERO for leaf in leaves})
    if not (rest.is_const() and float(rest.args[0]) == 0.0):
        raise ValueError("assign: the expression has a part that is not a Function")
    if not out:
        raise ValueError("assign: no Function in the expression")
    return out


def _argument(V, test):
    if V.nfields == 1:
        return U.Tensor(U.Scalar({((U.ZERO3, None) if test else (None, U.ZERO3)): S.ONE}))
    arr = np.empty((V.nfields,), dtype=object)
    for f in range(V.nfields):
        part = U.ZERO3 + (f,)                    # (a0, a1, a2, field)
        arr[f] = U.Scalar({((part, None) if test else (None, part)): S.ONE})
    return U.Tensor(arr)


def TrialFunction(V):
    """Scalar for a one-field space, a vector with one component per field otherwise
    (the MixedElement of common.py:337-351)."""
    return _argument(V, False)


def TestFunction(V):
    return _argument(V, True)


def derivative(form, u, du=None):
    """UFL ``derivative``: Gateaux derivative of ``form`` with respect to the Function
    ``u`` in the direction ``du`` -- a trial function by default
    (poisson-nonzero-bc.py:103), or the given TestFunction (first variation of an energy,
    kl-shell-svk/dynamic-tspline.py:232)."""
    if not isinstance(u, Function):
        raise TypeError("derivative() is taken with respect to a Function")
    test = False
    if du is not None:
        parts = [k for sc in U.as_tensor(du).a.ravel() for k in sc.terms]
        if not parts or any((k[0] is None) == (k[1] is None) for k in parts):
            raise ValueError("derivative(): du must be a TestFunction or a TrialFunction")
        test = parts[0][0] is not None
        if any((k[0] is not None) != test for k in parts):
            raise ValueError("derivative(): du mixes test and trial functions")
    if isinstance(u, VectorFunction):
        return U.gateaux(form, u.fid_fields(), test)
    return U.gateaux(form, u.fid, test)


def assemble(form, tensor=None):
    """dolfin.assemble on the FE space: float / FE vector / FE matrix."""
    owner = form.owner()
    if owner is None:
        return 0.0
    return owner._assemble_fe(form)


class File(object):
    """``File("x.pvd") << u`` writes the FE nodal values on the parametric
    mesh as a VTK structured grid (+ a .pvd index)."""

    def __init__(self, name):
        self.name = name
        self.count = 0

    def __lshift__(self, u):
        if isinstance(u, tuple):
            u = u[0]
        base, _ = os.path.splitext(self.name)
        d = os.path.dirname(self.name)
        if d:
            os.makedirs(d, exist_ok=True)
        patch = u.V.owner.patch()
        vals = dev.to_np(u.fe_tensor())
        X = patch.fe_node_coords()
        n = list(patch.nfe) + [1] * (3 - patch.dim)
        pts = np.zeros((X.shape[0], 3))
        pts[:, :patch.dim] = X
        fn = "%s%06d.vts" % (base, self.count)
        with open(fn, "w") as f:
            ext = "0 %d 0 %d 0 %d" % (n[0] - 1, n[1] - 1, n[2] - 1)
            f.write('<?xml version="1.0"?>\n<VTKFile type="StructuredGrid" version="0.1">\n')
            f.write('<StructuredGrid WholeExtent="%s"><Piece Extent="%s">\n' % (ext, ext))
            f.write('<PointData Scalars="u"><DataArray type="Float64" Name="u" format="ascii">\n')
            f.flush()          # ndarray.tofile writes through the descriptor, not Python's buffer
            np.asarray(vals, dtype=np.float64).tofile(f, sep=" ", format="%.17g")
            f.write('\n</DataArray></PointData>\n<Points><DataArray type="Float64" '
                    'NumberOfComponents="3" format="ascii">\n')
            f.flush()
            pts.ravel().tofile(f, sep=" ", format="%.17g")
            f.write('\n</DataArray></Points>\n</Piece></StructuredGrid>\n</VTKFile>\n')
        with open(self.name, "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n'
                    '<Collection>\n')
            for i in range(self.count + 1):
                f.write('<DataSet timestep="%d" part="0" file="%s%06d.vts" />\n'
                        % (i, os.path.basename(base), i))
            f.write('</Collection>\n</VTKFile>\n')
        self.count += 1
        return self


class KrylovSolver(object):
    """PETScKrylovSolver stand-in: carries tolerances for the device CG."""

    def __init__(self, method="cg", preconditioner="jacobi"):
        self.method = method
        self.preconditioner = preconditioner
        self.parameters = {"relative_tolerance": 1e-12, "absolute_tolerance": 0.0,
                           "maximum_iterations": 100000, "error_on_nonconvergence": False}


PETScKrylovSolver = KrylovSolver


# ---------------------------------------------------------------- generators
class AbstractExtractionGenerator(object):
    """common.py:130-502."""
    __metaclass__ = abc.ABCMeta

    def __init__(self, comm, *args):
        if not isinstance(comm, _Comm):
            args = (comm,) + args
            self.comm = worldcomm
        else:
            self.comm = comm
        self.customSetup(args)
        self.genericSetup()

    def getComm(self):
        return self.comm

    def useDG(self):
        return USE_DG_DEFAULT

    def extractionElement(self):
        return "DG" if self.useDG() else "Lagrange"

    def globalDof(self, field, localDof):
        off = 0
        for i in range(field):
            off += self.getNcp(i)
        return localDof + off

    def generatePermutation(self):
        return np.arange(sum(self.getNcp(i) for i in range(self.getNFields())))

    def addZeroDofsGlobal(self, newDofs):
        self.zeroDofs += newDofs

    def addZeroDofs(self, field, newDofs):
        off = self.globalDof(field, 0)
        if off == 0:
            self.addZeroDofsGlobal(newDofs if isinstance(newDofs, DofList) else list(newDofs))
        elif isinstance(newDofs, DofList):
            self.addZeroDofsGlobal(DofList.from_array(newDofs.asarray() + off))
        else:
            self.addZeroDofsGlobal([d + off for d in newDofs])

    def getPrealloc(self, control):
        return DEFAULT_PREALLOC

    def getIgnoreEps(self):
        return DEFAULT_BASIS_FUNC_IGNORE_EPS

    def genericSetup(self):
        """common.py:321-383; M / M_control are built on first access (they
        are never needed by the element-fused path)."""
        self.mesh = self.generateMesh()
        self.nsd = self.getNsd()
        fam = self.extractionElement()          # "DG" for discontinuous splines: the FE space is
        self.VE_control = (fam, self.getDegree(-1))   # then only a label (fused / matfree paths)
        self.VE = (fam, self.getDegree(0)) if self.getNFields() == 1 else \
            tuple((fam, self.getDegree(i)) for i in range(self.getNFields()))
        self.V_control = FunctionSpace(self, 1, control=True)
        self.V = FunctionSpace(self, self.getNFields())
        self._patch = None
        self._M = None
        self._M_control = None
        self._cpFuncs = None
        self.zeroDofs = DofList()

    # lazily built heavy objects ------------------------------------------
    def patch(self):
        raise NotImplementedError

    def M_matrix(self):
        return self.M

    @property
    def M(self):
        if self._M is None:
            self._M = self.generateM()
        return self._M

    @property
    def M_control(self):
        if self._M_control is None:
            self._M_control = self.generateM_control()
        return self._M_control

    @property
    def cpFuncs(self):
        if self._cpFuncs is None:
            P = self.controlNet()
            self._cpFuncs = []
            for i in range(self.nsd + 1):
                f = Function(self.V_control)
                f.set_iga(dev.from_np(P[:, i].copy()))
                self._cpFuncs.append(f)
        return self._cpFuncs

    def controlNet(self):
        """[ncp, nsd+1] homogeneous control points (loop of common.py:373-375)."""
        ncp = self.getNcp(-1)
        return np.array([[self.getHomogeneousCoordinate(I, i) for i in range(self.nsd + 1)]
                         for I in range(ncp)])

    def applyPermutation(self):
        pass        # single address space per GPU: IGA numbering is kept

    def writeExtraction(self, dirname, doPermutation=DEFAULT_DO_PERMUTATION):
        """Persist what ExtractedSpline(dirname, ...) needs (common.py:435-502):
        knots/degrees, control net and zero DoFs; same info-file grammar."""
        os.makedirs(dirname, exist_ok=True)
        with open(os.path.join(dirname, EXTRACTION_INFO_FILE), "w") as f:
            f.write("%d\n%s\n%d\n" % (self.nsd, self.extractionElement(), self.getNFields()))
            for i in range(-1, self.getNFields()):
                f.write("%d\n%d\n" % (self.getDegree(i), self.getNcp(i)))
        data = dict(P=self.controlNet(), zeroDofs=np.array(self.zeroDofs, dtype=np.int64))
        sp = self.getScalarSpline(-1) if hasattr(self, "getScalarSpline") else None
        if isinstance(sp, BSpline):
            data["degrees"] = np.array([s.p for s in sp.splines])
            for d, s in enumerate(sp.splines):
                data["knots%d" % d] = s.knots
        np.savez(os.path.join(dirname, EXTRACTION_DATA_FILE), **data)
        np.array(self.zeroDofs, dtype=np.int32).tofile(
            os.path.join(dirname, EXTRACTION_ZERO_DOFS_FILE))


class AbstractCoordinateChartSpline(AbstractExtractionGenerator):
    """common.py:1435-1669.  For tensor-product B-spline bases the per-node
    Python loop of generateM (:1497-1509) is replaced by the Kronecker-window
    kernel ``tg_m_fill``."""

    def isTensorProduct(self):
        """True when every field is a tensor-product ``BSpline`` (the structured fast path);
        False for a user ``AbstractScalarBasis`` (generic CSR path, tigar_b200/generic.py)."""
        return all(isinstance(self.getScalarSpline(i), BSpline)
                   for i in range(-1, self.getNFields()))

    def _tensor_spline(self, field):
        sp = self.getScalarSpline(field)
        if not isinstance(sp, BSpline):
            raise NotImplementedError("this operation needs a tensor-product BSpline basis")
        return sp

    def patch(self):
        if self._patch is None:
            if self.isTensorProduct():
                sp = self._tensor_spline(-1)
                self._patch = TensorPatch([s.p for s in sp.splines], None, splines=sp.splines,
                                          eps=self.getIgnoreEps())
            else:
                # generic basis: carrier B-splines bring its tensor-product extraction mesh
                # and the Q_p Lagrange tables; the IGA side is the basis' own M (CSR)
                from .bsplines import TensorMesh
                from .generic import carrier_splines
                if not isinstance(self.mesh, TensorMesh):
                    raise NotImplementedError(
                        "generic AbstractScalarBasis: generateMesh() must return a "
                        "tensor-product TensorMesh (element soups are out of scope)")
                if self.getNFields() != 1:
                    raise NotImplementedError("generic bases: one field")
                sp = carrier_splines(self.mesh, int(self.getDegree(-1)))
                self._patch = TensorPatch([s.p for s in sp], None, splines=sp,
                                          eps=self.getIgnoreEps())
                self._patch.generic = True
                self._patch.n_iga = int(self.getNcp(-1))
        return self._patch

    def generateM_control(self):
        if self.isTensorProduct():
            return self.patch().build_M()
        # the reference's loop (common.py:1497-1509): the user's basis evaluated at every FE node
        from .generic import CsrMatrix
        x = self.patch().fe_node_coords()
        rows = [self.getNodesAndEvals(x[I], -1) for I in range(x.shape[0])]
        return CsrMatrix.from_rows(rows, int(self.getNcp(-1)), self.getIgnoreEps())

    def equalOrder(self):
        """True if every field uses the control mesh's scalar spline (same degrees and
        knot vectors): the multi-field ``M`` (common.py:1546-1573) is then block
        diagonal with ``M_control`` in every block."""
        ctl = self.getScalarSpline(-1)
        for i in range(self.getNFields()):
            sp = self.getScalarSpline(i)
            if sp is ctl:
                continue
            if not (isinstance(sp, BSpline) and isinstance(ctl, BSpline)
                    and len(sp.splines) == len(ctl.splines)
                    and all(a.p == b.p and len(a.knots) == len(b.knots)
                            and np.array_equal(a.knots, b.knots)
                            for a, b in zip(sp.splines, ctl.splines))):
                return False
        return True

    def generateM(self):
        """Scalar extraction operator of one field block; the multi-field ``M`` of an
        equal-order spline is ``I_nFields (x) M_control`` and is never formed."""
        if self.equalOrder():
            return self.M_control
        raise NotImplementedError("fields of different order (FieldListSpline with "
                                  "unequal bases) are not built")

    def controlNet(self):
        cm = self.getControlMesh() if hasattr(self, "getControlMesh") else None
        if cm is not None and hasattr(cm, "controlNet"):
            return cm.controlNet()
        return AbstractExtractionGenerator.controlNet(self)


class AbstractMultiFieldSpline(AbstractCoordinateChartSpline):
    """common.py:1794-1885."""

    def getPrealloc(self, control):
        if control:
            return self.getScalarSpline(-1).getPrealloc()
        return max(self.getScalarSpline(i).getPrealloc() for i in range(self.getNFields()))

    def getScalarSpline(self, field):
        if field == -1:
            return self.getControlMesh().getScalarSpline()
        return self.getFieldSpline(field)

    def getNsd(self):
        return self.getControlMesh().getNsd()

    def getHomogeneousCoordinate(self, node, direction):
        return self.getControlMesh().getHomogeneousCoordinate(node, direction)

    def getNodesAndEvals(self, x, field):
        return self.getScalarSpline(field).getNodesAndEvals(x)

    def generateMesh(self):
        return self.getScalarSpline(-1).generateMesh(comm=self.comm)

    def getDegree(self, field):
        return self.getScalarSpline(field).getDegree()

    def getNcp(self, field):
        return self.getScalarSpline(field).getNcp()

    def useDG(self):
        return any(self.getScalarSpline(i).needsDG() for i in range(-1, self.getNFields()))


class EqualOrderSpline(AbstractMultiFieldSpline):
    """common.py:1891-1945."""

    def customSetup(self, args):
        self.numFields = args[0]
        self.controlMesh = args[1]

    def getNFields(self):
        return self.numFields

    def getControlMesh(self):
        return self.controlMesh

    def getFieldSpline(self, field):
        return self.getScalarSpline(-1)

    def addZeroDofsByLocation(self, subdomain, field):
        P = self.controlNet()
        nsd = self.getNsd()
        for I in range(P.shape[0]):
            x = P[I, :nsd] / P[I, nsd]
            if subdomain.inside(x, False) or subdomain.inside(x, True):
                self.zeroDofs += [self.globalDof(field, I)]


class FieldListSpline(AbstractMultiFieldSpline):
    """common.py:1948-1970."""

    def customSetup(self, args):
        self.controlMesh = args[0]
        self.fields = args[1]

    def getNFields(self):
        return len(self.fields)

    def getControlMesh(self):
        return self.controlMesh

    def getFieldSpline(self, field):
        return self.fields[field]


# ---------------------------------------------------------------- analysis
def _sorted_unique(z):
    """np.unique for an int64 index list (sort + adjacent compare; numpy's hash
    based path costs 30 ms on the 4e5 side DoFs of a 256^3 patch)."""
    z = np.sort(np.asarray(z, dtype=np.int64).ravel())
    if z.size == 0:
        return z
    keep = np.empty(z.size, dtype=bool)
    keep[0] = True
    np.not_equal(z[1:], z[:-1], out=keep[1:])
    return z[keep]


class ExtractedSpline(object):
    """common.py:667-1433."""
    _unit_weights = False        # all control weights exactly 1 (set by _set_control_net)

    def __init__(self, sourceArg, quadDeg, mesh=None, doPermutation=DEFAULT_DO_PERMUTATION,
                 comm=worldcomm, mode=None, controlNet=None):
        """``mode`` ("csr" / "fused") and ``controlNet`` (a ready [ncp, nsd+1]
        homogeneous control net, host array or list of device tensors, instead
        of the generator's per-node loop) are extensions of the reference
        signature (common.py:676-706)."""
        self._controlNetArg = controlNet
        if isinstance(sourceArg, AbstractExtractionGenerator):
            self.initFromGenerator(sourceArg, quadDeg, doPermutation)
        else:
            self.initFromFilesystem(sourceArg, quadDeg, comm, mesh)
        if self._patch.part is not None:
            if mode not in (None, "fused"):
                raise NotImplementedError("multi-GPU runs use the element-fused path")
            mode = "fused"
        if any(getattr(D, "discontinuous", False) for D in getattr(self._patch, "dirs", [])):
            # discontinuous splines (interior knot of multiplicity p+1): the reference extracts
            # to a DG space (BSplines.py:419-427); here only the paths that never form A_FE / M
            if mode == "csr":
                raise NotImplementedError("discontinuous B-splines: use mode='fused' or 'matfree' "
                                          "(the csr path would need a DG extraction space)")
            mode = mode or "fused"
        if getattr(self, "_generic", False):
            if mode not in (None, "csr"):
                raise NotImplementedError("a generic AbstractScalarBasis uses the csr path "
                                          "(FE assembly + M^T A M with its own M)")
            mode = "csr"
        self.mode = mode or os.environ.get("TIGAR_B200_MODE") or self._auto_mode()
        if self.mode not in ("csr", "fused", "matfree"):
            raise ValueError("mode must be 'csr', 'fused' or 'matfree'")
        if self.mode == "matfree" and (self._patch.part is not None or self.nFields != 1):
            raise NotImplementedError("matrix-free mode: one field, one GPU")
        self.genericSetup()

    # -- construction ------------------------------------------------------
    def initFromGenerator(self, generator, quadDeg, doPermutation=DEFAULT_DO_PERMUTATION):
        self.generator = generator
        self.quadDeg = quadDeg
        self.nsd = generator.getNsd()
        self.elementType = generator.extractionElement()
        self.nFields = generator.getNFields()
        self.p_control = generator.getDegree(-1)
        self.p = [generator.getDegree(i) for i in range(self.nFields)]
        self.mesh = generator.mesh
        self.comm = generator.getComm()
        self._generic = not generator.isTensorProduct()
        csize = self.comm.size
        part = (self.comm.rank, csize) if csize > 1 else None
        if self._generic:
            if part is not None:
                raise NotImplementedError("generic bases run on one GPU")
            base = generator.patch()                  # carrier of the FE side
            self._patch = TensorPatch(base.degrees, None, quadDeg=quadDeg, splines=base.splines,
                                      eps=generator.getIgnoreEps())
            self._patch.generic = True
            self._patch.n_iga = base.n_iga
        else:
            sp = generator._tensor_spline(-1)
            if self.nFields > 1 and not generator.equalOrder():
                raise NotImplementedError("multi-field splines are built for equal-order "
                                          "fields (every field on the control mesh's spline)")
            if self.nFields > 1 and part is not None:
                raise NotImplementedError("multi-field systems run on one GPU")
            self._patch = TensorPatch([s.p for s in sp.splines], None, quadDeg=quadDeg,
                                      splines=sp.splines, eps=generator.getIgnoreEps(),
                                      part=part)
        self.V = FunctionSpace(self, self.nFields)
        self.V_control = FunctionSpace(self, 1, control=True)
        self.VE, self.VE_control = generator.VE, generator.VE_control
        if self._controlNetArg is not None:
            P = self._controlNetArg
        else:
            cm = generator.getControlMesh() if hasattr(generator, "getControlMesh") else None
            if hasattr(cm, "controlNetDevice") and dev.device().type == "cuda":
                P = cm.controlNetDevice()          # generated on the device (no host loop)
            else:
                P = generator.controlNet()
        self._set_control_net(P)
        z = generator.zeroDofs
        # raw list (duplicates allowed, e.g. patch corners): the device mask does not need
        # it sorted; the public ``zeroDofs`` attribute is sorted and unique, made on demand
        self._zeroDofsRaw = z.asarray() if isinstance(z, DofList) else np.array(z, dtype=np.int64)
        self._zeroDofs = None
        self._M = None

    def initFromFilesystem(self, dirname, quadDeg, comm, mesh=None):
        """Reads what writeExtraction stored (common.py:748-894)."""
        from .bsplines import BSpline1
        data = np.load(os.path.join(dirname, EXTRACTION_DATA_FILE))
        with open(os.path.join(dirname, EXTRACTION_INFO_FILE)) as f:
            lines = f.read().split()
        self.generator = None
        self.quadDeg = quadDeg
        self.nsd = int(lines[0])
        self.elementType = lines[1]
        self.nFields = int(lines[2])
        self.p_control = int(lines[3])
        self.p = [int(lines[5 + 2 * i]) for i in range(self.nFields)]
        self.comm = comm
        if "degrees" not in data:
            raise NotImplementedError("stored extraction is not a tensor-product B-spline")
        deg = [int(x) for x in data["degrees"]]
        splines = [BSpline1(p, data["knots%d" % d]) for d, p in enumerate(deg)]
        self._patch = TensorPatch(deg, None, quadDeg=quadDeg, splines=splines)
        from .bsplines import TensorMesh
        self.mesh = TensorMesh([s.uniqueKnots for s in splines])
        self.V = FunctionSpace(self, self.nFields)
        self.V_control = FunctionSpace(self, 1, control=True)
        self.VE = self.VE_control = ("Lagrange", self.p_control)
        self._set_control_net(data["P"])
        self.zeroDofs = data["zeroDofs"].astype(np.int64)
        self._M = None

    @property
    def zeroDofs(self):
        """Sorted, unique global zero DoFs (int64)."""
        if self._zeroDofs is None:
            self._zeroDofs = _sorted_unique(self._zeroDofsRaw)
        return self._zeroDofs

    @zeroDofs.setter
    def zeroDofs(self, z):
        self._zeroDofsRaw = np.asarray(z, dtype=np.int64).ravel()
        self._zeroDofs = None
        self._mask = None

    def _set_control_net(self, P):
        """P: [ncp, nsd+1] host array (torch pinned or numpy), or a list of
        nsd+1 device tensors (already resident)."""
        import torch
        self.cpFuncs = []
        unit = getattr(P, "unit_weights", None)
        if isinstance(P, (list, tuple)):
            cols = list(P)
            self.controlNet = None
        else:
            t = P if isinstance(P, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(P, dtype=np.float64))
            self.controlNet = t
            patch = self._patch
            if getattr(patch, "part", None) is not None:
                # slab partition: upload only the planes this rank's cell layers read
                from .engine import SlabVector
                p0, p1 = patch.coef_planes()
                pl = patch.plane
                d = t[p0 * pl:p1 * pl].to(dev.device(), non_blocking=True)
                cols = [SlabVector(d[:, i].contiguous(), p0, p1, pl)
                        for i in range(self.nsd + 1)]
                self._h2d_bytes = d.numel() * 8
                if unit is None:
                    unit = os.environ.get("TIGAR_B200_UNIT_WEIGHTS", "1") == "1" and \
                        bool((cols[self.nsd].t == 1.0).all().item())
            else:
                d = t.to(dev.device(), non_blocking=True)        # one H2D copy
                cols = [d[:, i].contiguous() for i in range(self.nsd + 1)]
                self._h2d_bytes = d.numel() * 8
        if unit is None:
            unit = os.environ.get("TIGAR_B200_UNIT_WEIGHTS", "1") == "1" and \
                bool((cols[self.nsd] == 1.0).all().item())
        self._unit_weights = bool(unit)
        for i in range(self.nsd + 1):
            f = Function(self.V_control)
            f.set_iga(cols[i])
            self.cpFuncs.append(f)

    def patch(self):
        return self._patch

    def _auto_mode(self):
        """'csr' while the global operands fit comfortably in HBM."""
        import torch
        p = self._patch
        need = 8 * (p.window("A").nnz + p.window("M").nnz + p.window("P").nnz
                    + p.window("C").nnz) + 8 * (2 * p.n_fe)
        free, _ = torch.cuda.mem_get_info()
        if need < 0.5 * free:
            return "csr"
        # the element-fused path still stores C; beyond that only the operator action fits
        if self.nFields == 1 and p.part is None and 8 * p.window("C").nnz > 0.9 * free:
            return "matfree"
        return "fused"

    def M_matrix(self):
        if self._M is None:
            if getattr(self, "_generic", False):
                self._M = self.generator.M            # the basis' own rows (CSR)
            elif self.generator is not None and self.generator._M is not None:
                self._M = self.generator._M
            else:
                self._M = self._patch.build_M()
        return self._M

    @property
    def M(self):
        return self.M_matrix()

    @property
    def M_control(self):
        return self.M_matrix()

    def genericSetup(self):
        """Symbolic geometry, common.py:896-966 + calculusUtils.py."""
        dim = self._patch.dim
        U.DEFAULT_DIM[0] = dim
        self.boundaryMarkers = None
        if self._unit_weights:
            # all weights are exactly 1 (explicit B-spline meshes, BSplines.py:935-960): the
            # weight function is the partition of unity, F = P/w = P up to one rounding
            comps = [self.cpFuncs[i] for i in range(self.nsd)]
        else:
            comps = [self.cpFuncs[i] / self.cpFuncs[self.nsd] for i in range(self.nsd)]
        self.F = U.as_vector(comps)
        self.DF = U.parametric_grad(self.F, dim)                    # [nsd, dim]
        self.g = U.dot(self.DF.T, self.DF)                          # getMetric
        self.N = None
        self.n = None
        if self.nsd == dim and os.environ.get("TIGAR_B200_SQUARE_MAP", "1") == "1":
            # square map (volume patch in 3-D, planar patch with nsd = 2): sqrt(det(DF^T DF)) =
            # |det DF| and (DF^T DF)^-1 DF^T = DF^-1 -- the same quantities as
            # calculusUtils.py:18-24, 56-69 in half the operations per Gauss point
            J = U.abs_(U.det(self.DF))
            self.pinvDF = U.inv(self.DF)
        else:
            J = U.sqrt(U.det(self.g))                               # volumeJacobian
            self.pinvDF = U.dot(U.inv(self.g), self.DF.T)           # pinvD
        self.dx = U.Measure(J, self, "dx")
        self.ds = U.Measure(None, self, "ds")
        self.gamma = None
        self.setSolverOptions()
        self._mask = None

    # -- differential operators (calculusUtils.py:255-276) -----------------
    def grad(self, f, F=None):
        if F is not None:
            raise NotImplementedError("grad with an alternative mapping")
        return U.dot(U.parametric_grad(f, self._patch.dim), self.pinvDF)

    def div(self, f, F=None):
        g = self.grad(f, F)
        n = g.a.ndim
        if n < 2:
            raise ValueError("div of a scalar")
        out = np.empty(g.a.shape[:-2], dtype=object)
        for i in np.ndindex(*g.a.shape[:-2]):
            tot = U.Scalar()
            for k in range(g.a.shape[-1]):
                tot = tot.add(g.a[i + (k, k)])
            out[i] = tot
        return U.Tensor(out) if out.shape else U.Tensor(out[()])

    def curl(self, f, F=None):
        """Physical curl (calculusUtils.py:278-302): vector in 3-D, scalar for a
        2-D vector, vector for a 2-D scalar."""
        return U.curl_from_grad(f, self.grad(f, F))

    def GRAD(self, f):
        """Covariant derivative w.r.t. the parametric chart with the new index raised
        (common.py:1068-1080); a plain tensor is taken with all indices lowered."""
        from . import calculus as CU
        ff = f if isinstance(f, CU.CurvilinearTensor) else CU.CurvilinearTensor(f, self.g)
        return CU.curvilinearGrad(ff)

    def DIV(self, f):
        """Curvilinear divergence matching GRAD (common.py:1081-1093); a plain tensor
        is taken with all indices raised."""
        from . import calculus as CU
        ff = f if isinstance(f, CU.CurvilinearTensor) else CU.CurvilinearTensor(f, self.g).sharp()
        return CU.curvilinearDiv(ff)

    def parametricExpression(self, expr):
        """``Expression`` in the parametric coordinates (common.py:1111-1117); the
        formula is evaluated exactly at the Gauss points, not interpolated."""
        return U.expression_from_string(expr, self.parametricCoordinates())

    def parametricGrad(self, f):
        return U.parametric_grad(f, self._patch.dim)

    def parametricCoordinates(self):
        return U.as_vector([U.Tensor(U.Scalar.coef(S.xi(d))) for d in range(self._patch.dim)])

    def spatialCoordinates(self):
        return self.F

    def rationalize(self, u):
        if self._unit_weights:
            return u
        return u / self.cpFuncs[self.nsd]

    # -- assembly ------------------------------------------------------------
    def _funcs(self, kind):
        out = {}
        for fid, f in list(_functions.items()):
            if f.V.owner is not self and f.V.owner is not self.generator:
                continue
            if kind == "iga":
                if f.iga is not None:
                    out[fid] = f.iga
                elif f._fe is None and not f.V.control:
                    out[fid] = _LazyZero(f, self._patch.n_iga)   # never assigned: zero (dolfin)
                elif f._fe is not None and self.nFields == 1:
                    out[fid] = _LazyIGA(f, self)     # only FE data: derive the IGA DoFs
            else:
                out[fid] = _LazyFE(f)
        return _FuncTable(out)

    def _weighted(self, scalar):
        """Multiply every coefficient by the quadrature weight register."""
        w = S.wq()
        return {k: S.mul(v, w) for k, v in scalar.terms.items()}

    def n_total(self):
        """Number of IGA DoFs of all fields."""
        return self._patch.n_iga * self.nFields

    def _assemble_blocks(self, terms, ar, kind):
        """Multi-field: one scalar assembly per (test field, trial field) block."""
        from . import multifield as MF
        p, nf = self._patch, self.nFields
        funcs = self._funcs(kind)
        if ar == 2:
            blocks = {}
            for fg, bt in sorted(MF.split_matrix_terms(terms, nf).items()):
                blocks[fg] = p.assemble_matrix(bt, funcs, kind,
                                               cache=self._program_cache(("m", kind, fg), bt))
            W = p.window("A" if kind == "fe" else "C")
            return MF.BlockMatrix(nf, blocks, W.nrows)
        parts = MF.split_vector_terms(terms, nf)
        n = p.n_fe if kind == "fe" else p.n_iga
        out = dev.zeros(nf * n)
        for f, vt in sorted(parts.items()):
            p.assemble_vector(vt, funcs, kind, out=out[f * n:(f + 1) * n],
                              cache=self._program_cache(("v", kind, f), vt))
        return out

    def _program_cache(self, tag, terms):
        """Per-block cache of compiled Gauss-point programs for Newton / time loops that
        assemble the SAME forms again and again (engine._setup_cached; symbolic compilation
        of a shell tangent block costs ~0.2 s of Python per call).  Keyed by the identity
        of the (hash-consed, immutable) coefficient nodes and the current values of the
        mutable Parameters, which are baked into a program when it is compiled.
        Opt-in (``TIGAR_B200_PROG_CACHE=1``) until it has run on a device."""
        if os.environ.get("TIGAR_B200_PROG_CACHE") != "1":
            return None
        if getattr(self, "_prog_cache", None) is None:
            self._prog_cache = {}
        key = (tag, tuple(sorted((k, n.uid) for k, n in terms.items())),
               tuple(sorted(S.PARAMS.items())))
        return self._prog_cache.setdefault(key, {})

    def _assemble_kind(self, form, kind):
        sc = form.scalar()
        ar = sc.arity()
        terms = self._weighted(sc)
        p = self._patch
        if self.nFields > 1 and ar > 0:
            return self._assemble_blocks(terms, ar, kind)
        if ar == 2:
            return p.assemble_matrix({(k[0], k[1]): v for k, v in terms.items()},
                                     self._funcs(kind), kind)
        if ar == 1:
            if any(k[0] is None for k in terms):
                raise ValueError("linear form must be linear in the TEST function")
            return p.assemble_vector({k[0]: v for k, v in terms.items()}, self._funcs(kind), kind)
        return p.assemble_scalar(terms.get((None, None), S.ZERO), self._funcs(kind), kind)

    def _assemble_fe(self, form):
        """dolfin.assemble(form): FE-space tensors (floats for functionals)."""
        ar = form.arity()
        if ar == 0:
            return self._assemble_kind(form, "fe" if self.mode == "csr" else "iga")
        r = self._assemble_kind(form, "fe")
        return DeviceVector(r) if ar == 1 else r

    def _bc_mask(self):
        """Device 0/1 mask over the IGA DoFs of all fields (field-major)."""
        if self._mask is None:
            if self.nFields == 1:
                self._mask = self._patch.bc_mask(self._zeroDofsRaw)
            else:
                from . import multifield as MF
                import torch
                per = MF.split_zero_dofs(self.zeroDofs, self.nFields, self._patch.n_iga)
                self._mask = torch.cat([self._patch.bc_mask(z) for z in per])
        return self._mask

    def _field_mask(self, f):
        n = self._patch.n_iga
        return self._bc_mask()[f * n:(f + 1) * n]

    def _apply_bcs_blocks(self, Cm, diag):
        """zeroRowsColumns(zeroDofs, diag) (common.py:1199-1200) on the block grid: rows
        of the test field's zero DoFs, columns of the trial field's, ``diag`` on the
        diagonal of the diagonal blocks (which must exist to carry it)."""
        p = self._patch
        for f in range(self.nFields):
            if Cm.block(f, f) is None:
                Cm.blocks[(f, f)] = WinMatrix(p.window("C"))
        for (f, g), B in Cm.blocks.items():
            check(lib.tg_win_zero_rows_cols(B.window.ref(), dev.ptr(B.vals),
                                            dev.ptr(self._field_mask(f)),
                                            dev.ptr(self._field_mask(g)),
                                            float(diag) if f == g else 0.0, 0, dev.stream()))
        return Cm

    def extractVector(self, b, applyBCs=True):
        """M^T b (+ zero BC entries), common.py:1142-1160."""
        t = b.t if isinstance(b, DeviceVector) else b
        if self.nFields > 1:
            nfe, n = self._patch.n_fe, self._patch.n_iga
            MTb = dev.empty(self.nFields * n)
            for f in range(self.nFields):
                MTb[f * n:(f + 1) * n].copy_(
                    self._patch.mt_vec(self.M_matrix(), t[f * nfe:(f + 1) * nfe].contiguous()))
        else:
            MTb = self._patch.mt_vec(self.M_matrix(), t)
        if applyBCs:
            self._patch.apply_bcs_vector(MTb, self._bc_mask())
        return DeviceVector(MTb)

    def assembleVector(self, form, applyBCs=True):
        if self.mode == "csr":
            return self.extractVector(self._assemble_kind(form, "fe"), applyBCs)
        MTb = self._assemble_kind(form, "iga")
        if applyBCs:
            self._patch.apply_bcs_vector(MTb, self._bc_mask())
        return DeviceVector(MTb, distributed=self._patch.part is not None)

    def extractMatrix(self, A, applyBCs=True, diag=1):
        """M^T A M then zeroRowsColumns, common.py:1176-1204."""
        if self.nFields > 1:
            from . import multifield as MF
            MTAM = MF.BlockMatrix(self.nFields,
                                  {fg: self._patch.ptap(B, self.M_matrix())
                                   for fg, B in sorted(A.blocks.items())}, self._patch.n_iga)
            return self._apply_bcs_blocks(MTAM, diag) if applyBCs else MTAM
        MTAM = self._patch.ptap(A, self.M_matrix())
        if applyBCs:
            self._patch.apply_bcs_matrix(MTAM, self._bc_mask(), diag)
        return MTAM

    def assembleMatrix(self, form, applyBCs=True, diag=1):
        if self.mode == "matfree":
            from .matfree import FormOperator
            sc = form.scalar()
            if sc.arity() != 2:
                raise ValueError("assembleMatrix needs a bilinear form")
            return FormOperator(self, {(k[0], k[1]): v for k, v in self._weighted(sc).items()},
                                applyBCs, diag)
        if self.mode == "csr":
            return self.extractMatrix(self._assemble_kind(form, "fe"), applyBCs, diag)
        MTAM = self._assemble_kind(form, "iga")
        if applyBCs:
            if self.nFields > 1:
                return self._apply_bcs_blocks(MTAM, diag)
            self._patch.apply_bcs_matrix(MTAM, self._bc_mask(), diag)
        return MTAM

    def assembleLinearSystem(self, lhsForm, rhsForm, applyBCs=True):
        """common.py:1223-1234.  Matrix and vector share one Gauss-point pass."""
        kind = "fe" if self.mode == "csr" else "iga"
        ms, vs = lhsForm.scalar(), rhsForm.scalar()
        if (self.mode == "matfree" and self.nFields == 1 and ms.arity() == 2 and vs.arity() == 1
                and not any(k[0] is None for k in vs.terms)
                and os.environ.get("TIGAR_B200_MF_DIAG_PASS", "1") == "1"):
            # load vector and Jacobi diagonal of the operator from one Gauss-point pass
            op = self.assembleMatrix(lhsForm, applyBCs)
            vt = {k[0]: n for k, n in self._weighted(vs).items()}
            both = getattr(self._patch, "assemble_vector_and_diag", None)
            r = both(vt, op.mterms, self._funcs("iga")) if both is not None else None
            if r is not None:
                op.set_diagonal(r[1])
                if applyBCs:
                    self._patch.apply_bcs_vector(r[0], self._bc_mask())
                return op, DeviceVector(r[0])
            return op, self.assembleVector(rhsForm, applyBCs)
        if (self.nFields > 1 or self.mode == "matfree" or ms.arity() != 2 or vs.arity() != 1
                or any(k[0] is None for k in vs.terms)):
            return (self.assembleMatrix(lhsForm, applyBCs),
                    self.assembleVector(rhsForm, applyBCs))
        mt = {(k[0], k[1]): n for k, n in self._weighted(ms).items()}
        vt = {k[0]: n for k, n in self._weighted(vs).items()}
        A, b = self._patch.assemble_system(mt, vt, self._funcs(kind), kind)
        if self.mode == "csr":
            return (self.extractMatrix(A, applyBCs), self.extractVector(b, applyBCs))
        if applyBCs:
            self._patch.apply_bcs_matrix(A, self._bc_mask(), 1)
            self._patch.apply_bcs_vector(b, self._bc_mask())
        return A, DeviceVector(b, distributed=self._patch.part is not None)

    def solveLinearSystem(self, MTAM, MTb, u):
        """common.py:1236-1263: returns the IGA DoF vector, updates ``u``."""
        ls = self.linearSolver
        prm = ls.parameters if ls is not None else {}
        rtol = prm.get("relative_tolerance", self.cgRelativeTolerance)
        atol = prm.get("absolute_tolerance", 0.0)
        maxit = prm.get("maximum_iterations", 200000)
        from .matfree import FormOperator
        from .generic import GenericPtAP
        if isinstance(MTAM, GenericPtAP):
            x0 = None if u.iga is None else u.iga.clone()
            x, its, rel = MTAM.solve(MTb.t, x0, rtol, atol, maxit)
            self.lastSolve = dict(iterations=its, relative_residual=rel, method="cg")
            self._check_converged(ls, its, rel, maxit, rtol, atol)
            u.set_iga(x)
            return DeviceVector(x)
        if isinstance(MTAM, FormOperator):
            from .matfree import solve_matfree_cg, solve_matfree_fd
            method = os.environ.get("TIGAR_B200_SOLVER", self._solver_method(ls))
            if method == "jacobi":
                x, its, rel = solve_matfree_cg(MTAM, MTb.t, rtol, atol, maxit)
            else:
                x, its, rel = solve_matfree_fd(MTAM, MTb.t, rtol, atol, maxit)
                method = "fd"
            self.lastSolve = dict(iterations=its, relative_residual=rel, method=method)
            self._check_converged(ls, its, rel, maxit, rtol, atol)
            u.set_iga(x)
            return DeviceVector(x)
        if self.nFields > 1:
            from . import multifield as MF
            x, its, rel, used = MF.solve_block(MTAM, MTb.t, rtol, atol, maxit,
                                               self._solver_method(ls))
            self.lastSolve = dict(iterations=its, relative_residual=rel, method=used)
            self._check_converged(ls, its, rel, maxit, rtol, atol)
            u.set_iga(x)
            return DeviceVector(x)
        x0 = None if u.iga is None else u.iga.clone()
        method = self._solver_method(ls)
        mask = getattr(MTAM, "bc_mask", None)
        x, its, rel, used = self._patch.solve(MTAM, MTb.t, x0, rtol, atol, maxit, method, mask,
                                              getattr(MTAM, "bc_diag", 1.0))
        self.lastSolve = dict(iterations=its, relative_residual=rel, method=used)
        self._check_converged(ls, its, rel, maxit, rtol, atol)
        if self._patch.part is not None:       # replicate the IGA DoF vector on every rank
            from .multigpu import gather_planes
            x = gather_planes(x, self._patch)
        u.set_iga(x)
        return DeviceVector(x)

    @staticmethod
    def _solver_method(ls):
        """Which device solver stands in for ``dolfin.solve``: no ``linearSolver`` means the
        reference's default direct LU (common.py:1255-1256) -> "auto" (band Cholesky while it
        fits, FD-preconditioned CG beyond); a user Krylov solver with a Jacobi / no
        preconditioner gets Jacobi-CG; any stronger PETSc preconditioner name maps to FD-CG."""
        if ls is None:
            return "auto"
        m = getattr(ls, "solver_method", None)
        if m:
            return m
        pc = str(getattr(ls, "preconditioner", "jacobi")).lower()
        if pc in ("jacobi", "none", "default", "bjacobi"):
            return "jacobi"
        if pc in ("lu", "cholesky", "direct"):
            return "direct"
        return "fd"

    def _check_converged(self, ls, its, rel, maxit, rtol, atol):
        """The reference's LU cannot 'not converge'; an iterative stand-in can.  Raise (or
        warn, when the user's solver says error_on_nonconvergence=False) instead of returning
        an unconverged vector silently."""
        if its < maxit or rel <= max(rtol, 1e-300):
            return
        msg = ("linear solver stopped at the iteration limit (%d) with relative residual %.3e "
               "> %.3e" % (its, rel, rtol))
        prm = ls.parameters if ls is not None else {}
        if prm.get("error_on_nonconvergence", True):
            raise RuntimeError(msg)
        import warnings
        warnings.warn(msg)

    def solveLinearVariationalProblem(self, residualForm, u, applyBCs=True):
        if isinstance(residualForm, U.Equation):
            lhsForm, rhsForm = residualForm.lhs, residualForm.rhs
        else:
            lhsForm, rhsForm = U.lhs(residualForm), U.rhs(residualForm)
        if rhsForm.empty():          # common.py:1285-1287: zero right-hand side
            MTAM = self.assembleMatrix(lhsForm, applyBCs)
            MTb = DeviceVector(dev.zeros(self.n_total()))
        else:
            MTAM, MTb = self.assembleLinearSystem(lhsForm, rhsForm, applyBCs)
        return self.solveLinearSystem(MTAM, MTb, u)

    def setSolverOptions(self, maxIters=20, relativeTolerance=1e-5, linearSolver=None):
        self.maxIters = maxIters
        self.relativeTolerance = relativeTolerance
        self.linearSolver = linearSolver
        # DOLFIN's default solve() is a direct LU (common.py:1255-1256); the
        # device CG is run to a residual that makes the difference invisible
        # at the 1e-10 parity bar.
        self.cgRelativeTolerance = float(os.environ.get("TIGAR_B200_CG_RTOL", "1e-13"))

    def solveNonlinearVariationalProblem(self, residualForm, J, u, referenceError=None,
                                         igaDoFs=None):
        """Newton iteration of common.py:1304-1348."""
        if igaDoFs is not None:
            u.set_iga(igaDoFs.t.clone())
        # a Function that was never assigned is zero (dolfin); give it IGA data so the
        # element-fused path can read it
        for c in (u.comps if isinstance(u, VectorFunction) else [u]):
            if c.iga is None and c._fe is None:
                c.set_iga(dev.zeros(self._patch.n_iga))
        converged = False
        for i in range(self.maxIters):
            MTAM, MTb = self.assembleLinearSystem(J, residualForm)
            currentNorm = norm(MTb)
            if i == 0 and referenceError is None:
                referenceError = currentNorm
            relativeNorm = currentNorm / referenceError
            if MPI.rank(self.comm) == 0:
                print("Solver iteration: " + str(i) + " , Relative norm: " + str(relativeNorm))
                sys.stdout.flush()
            if relativeNorm < self.relativeTolerance:
                converged = True
                break
            du = Function(self.V)
            inc = self.solveLinearSystem(MTAM, MTb, du)
            base = u.iga if u.iga is not None else dev.zeros(self.n_total())
            u.set_iga(base - inc.t)
            if igaDoFs is not None:
                igaDoFs.t.copy_(u.iga)
        if not converged:
            print("ERROR: Nonlinear solver failed to converge.")
            raise SystemExit(1)

    def project(self, toProject, applyBCs=False, rationalize=True, lumpMass=False):
        """L2 projection onto the spline space, common.py:1392-1433."""
        u = self.rationalize(TrialFunction(self.V))
        v = self.rationalize(TestFunction(self.V))
        retval = Function(self.V)
        if not lumpMass:
            self.solveLinearVariationalProblem(
                U.inner(u, v) * self.dx == U.inner(toProject, v) * self.dx, retval, applyBCs)
        else:
            # row-sum lumping (common.py:1416-1430): U = (M^T b) / (M^T m), m = int 1 . v
            one = U.Constant(1.0) if self.nFields == 1 else U.Constant(self.nFields * (1.0,))
            lhs = self.assembleVector(U.inner(one, v) * self.dx, applyBCs=False)
            rhs = self.assembleVector(U.inner(toProject, v) * self.dx, applyBCs=applyBCs)
            retval.set_iga(rhs.t / lhs.t)
        return self.rationalize(retval) if rationalize else retval

    def FEtoIGA(self, u):
        """Testing helper of the reference (common.py:968-993): IGA DoFs from the FE
        coefficients of ``u`` by the pseudo-inverse problem (M^T M) U = M^T u, with
        M^T M formed as the triple product M^T I M on the FE pattern (as inefficient as
        the reference says it is) and solved by the device CG.  Returns the IGA vector."""
        p, M = self._patch, self.M_matrix()
        n_fe, n = p.n_fe, p.n_iga
        ident = WinMatrix(p.window("A"))
        ones = dev.zeros(n_fe, dev.U8) + 1
        check(lib.tg_win_zero_rows_cols(ident.window.ref(), dev.ptr(ident.vals), dev.ptr(ones),
                                        dev.ptr(ones), 1.0, 0, dev.stream()))
        MTM = p.ptap(ident, M)
        del ident
        fe = u.fe_tensor()
        out = dev.empty(self.nFields * n)
        for f in range(self.nFields):
            rhs = p.mt_vec(M, fe[f * n_fe:(f + 1) * n_fe].contiguous())
            x, its, rel = p.solve_cg(MTM, rhs, None, self.cgRelativeTolerance, 0.0, 200000)
            out[f * n:(f + 1) * n].copy_(x)
        return DeviceVector(out)


class _LazyFE(object):
    """Defers M*U until a kernel really needs the FE coefficients."""

    def __init__(self, f):
        self.f = f

    def resolve(self):
        return self.f.fe_tensor()


class _LazyIGA(object):
    """A Function that only has FE coefficients (written through ``u.vector()``): its IGA DoFs
    are the pseudo-inverse problem of FEtoIGA (common.py:968-993), solved when first needed."""

    def __init__(self, f, spline):
        self.f, self.spline = f, spline

    def resolve(self):
        if self.f.iga is None:
            fe = self.f._fe
            self.f.iga = self.spline.FEtoIGA(self.f).t
            self.f._fe = fe
        return self.f.iga


class _LazyZero(object):
    """A Function that was never assigned reads as zero; its IGA data is created the first
    time a kernel needs it."""

    def __init__(self, f, n):
        self.f, self.n = f, n

    def resolve(self):
        if self.f.iga is None:
            self.f.set_iga(dev.zeros(self.n))
        return self.f.iga


class _FuncTable(dict):
    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if isinstance(v, (_LazyFE, _LazyZero, _LazyIGA)):
            v = v.resolve()
            dict.__setitem__(self, k, v)
        return v
