"""Synthetic workloads that drive the hot path the way OpenCMP's solver loop does.

``INSTaylorGreen`` restates, against the NGSolve-style front end, exactly what the reference executes per time step
for ``examples/INS`` (HDiv-DG order k / L2 order k-1, Oseen linearisation) with the implicit-Euler scheme:

  forms      opencmp/models/ins.py:178-321 (construct_bilinear_time_ODE / _time_coefficient / construct_linear)
  scheme     opencmp/solvers/time_integration_schemes.py:79-132 (implicit_euler) + base_model.py:418-497 (mass terms)
  time step  opencmp/solvers/base_solver.py:521-587 (_solve: advance t, apply Dirichlet BCs, re-assemble, solve)
  Picard     opencmp/models/ins.py:323-355 (assemble, precondition, linear_solve, ||W-u||_L2, W <- u)
  solve      opencmp/models/base_model.py:886-947 (linear_solve dispatch)

on a structured [0,pi]^2 triangle mesh with the Taylor-Green data of examples/INS/{bc_dir,ic_dir,ref_sol_dir}.
"""
from __future__ import annotations

import numpy as np

from . import ngs
from .mesh import structured_2d


def _jump(q):
    return q - q.Other()


def _avg(q):
    return 0.5 * (q + q.Other())


def _grad_avg(q):
    return 0.5 * (ngs.Grad(q) + ngs.Grad(q.Other()))


class INSTaylorGreen:
    def __init__(self, N: int, order: int = 3, dt: float = 1e-3, nu: float = 1.0, ipc: float = 10.0,
                 linear_solver: str = 'GMRes', preconditioner: str = 'multigrid', linear_tolerance: float = 1e-10,
                 linear_max_iterations: int = 500, nonlinear_max_iterations: int = 3,
                 nonlinear_tolerance=(1e-4, 1e-6), mesh=None):
        if mesh is None and preconditioner == 'multigrid':
            # nested hierarchy: uniform (red) refinement of a coarse structured mesh, kept by Mesh.Refine()
            n0 = N
            import os
            nmin = int(os.environ.get('OCMP_MG_N0', '4'))
            while n0 % 2 == 0 and n0 > nmin:
                n0 //= 2
            mesh = structured_2d([n0, n0], scale=(np.pi, np.pi))
            while n0 < N:
                mesh.Refine()
                n0 *= 2
        self.mesh = ngs.Mesh(mesh if mesh is not None else structured_2d([N, N], scale=(np.pi, np.pi)))
        m = self.mesh
        k = order
        self.order = k
        self.linear_solver, self.preconditioner = linear_solver, preconditioner
        self.linear_tolerance, self.linear_max_iterations = linear_tolerance, linear_max_iterations
        self.nonlinear_max_iters = nonlinear_max_iterations
        self.rel_nonlinear_tolerance, self.abs_nonlinear_tolerance = nonlinear_tolerance
        dnames = 'top|bottom|left|right'
        self.dirichlet = dnames
        # models/ins.py:95-128
        self.V = ngs.HDiv(m, order=k, dirichlet=dnames, dgjumps=True)
        self.Q = ngs.L2(m, order=k - 1, dgjumps=True)
        self.fes = ngs.FESpace([self.V, self.Q], dgjumps=True)
        (u, p), (v, q) = self.fes.TrialFunction(), self.fes.TestFunction()
        self.t = ngs.Parameter(0.0)
        self.dt = ngs.Parameter(dt)
        t = self.t
        x, y = ngs.x, ngs.y
        kv = ngs.CoefficientFunction(nu)
        self.u_ref = ngs.CoefficientFunction((-ngs.cos(x) * ngs.sin(y) * ngs.exp(-2 * nu * t),
                                              ngs.sin(x) * ngs.cos(y) * ngs.exp(-2 * nu * t)))
        self.p_ref = -0.25 * (ngs.cos(2 * x) + ngs.cos(2 * y)) * ngs.exp(-4 * nu * t)
        g = self.u_ref
        f = ngs.CoefficientFunction((0.0, 0.0))
        n = ngs.specialcf.normal(2)
        h = ngs.specialcf.mesh_size
        alpha = (ipc * k ** 2) / h                       # base_model.py:150, helpers/ngsolve_.py:67
        self.gfu = ngs.GridFunction(self.fes)
        self.gfu_0 = ngs.GridFunction(self.fes)          # u^n
        self.W = ngs.GridFunction(self.V)                # Oseen wind, models/ins.py:130-134
        w = self.W
        dtp = self.dt
        ds_d = ngs.ds(skeleton=True, definedon=m.Boundaries(dnames))
        dS = ngs.dx(skeleton=True)
        a = ngs.BilinearForm(self.fes)
        # construct_bilinear_time_coefficient, ins.py:228-270
        a += dtp * (-ngs.div(u) * q - ngs.div(v) * p - 1e-10 * p * q) * ngs.dx
        ju, jv = _jump(u), _jump(v)
        a += dtp * (kv * alpha * ngs.InnerProduct(ju, jv)
                    - kv * ngs.InnerProduct(_grad_avg(u), ngs.OuterProduct(jv, n))
                    - kv * ngs.InnerProduct(_grad_avg(v), ngs.OuterProduct(ju, n))) * dS
        a += dtp * jv * (w * n * _avg(u) + 0.5 * ngs.Norm(w * n) * ju) * dS
        # construct_bilinear_time_ODE, ins.py:178-226
        a += dtp * (kv * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v))) * ngs.dx
        a += -dtp * ngs.InnerProduct(ngs.OuterProduct(u, w), ngs.Grad(v)) * ngs.dx
        a += dtp * (kv * alpha * u * v - kv * ngs.InnerProduct(ngs.Grad(u), ngs.OuterProduct(v, n))
                    - kv * ngs.InnerProduct(ngs.Grad(v), ngs.OuterProduct(u, n))) * ds_d
        a += dtp * v * (0.5 * w * n * u + 0.5 * ngs.Norm(w * n) * u) * ds_d
        # time derivative, base_model.py:469-471
        a += (u * v) * ngs.dx
        L = ngs.LinearForm(self.fes)
        L += dtp * v * f * ngs.dx
        L += dtp * (kv * alpha * g * v - kv * ngs.InnerProduct(ngs.Grad(v), ngs.OuterProduct(g, n))) * ds_d
        L += dtp * v * (-0.5 * w * n * g + 0.5 * ngs.Norm(w * n) * g) * ds_d
        L += (self.gfu_0.components[0] * v) * ngs.dx
        self.a, self.L = a, L
        self.pre = ngs.Preconditioner(a, preconditioner) if preconditioner is not None else None
        # initial condition (examples/INS/ic_dir/ic_config) and wind
        self.gfu.components[0].Set(self.u_ref)
        self.gfu.components[1].Set(self.p_ref)
        self.gfu_0.vec.data = self.gfu.vec
        self.W.vec.data = self.gfu.components[0].vec
        self.picard_iterations = 0
        self.linear_iterations = []

    @property
    def ndof(self) -> int:
        return self.fes.ndof

    @property
    def nnz(self) -> int:
        return self.fes.pattern().nnz

    def apply_dirichlet_bcs(self):
        """base_model.py:321-341"""
        self.gfu.components[0].Set(self.u_ref, definedon=self.mesh.Boundaries(self.dirichlet))

    def assemble(self):
        """base_solver.py:368-377"""
        self.a.Assemble()
        self.L.Assemble()
        if self.pre is not None:
            self.pre.Update()

    def linear_solve(self):
        """base_model.py:886-947"""
        be = ngs.get_backend()
        if self.linear_solver == 'direct':
            inv = self.a.mat.Inverse(self.fes.FreeDofs())
            r = self.L.vec.CreateVector()
            r.data = self.L.vec - self.a.mat * self.gfu.vec
            self.gfu.vec.data += inv * r
        elif self.linear_solver == 'GMRes':
            ngs.solvers.GMRes(A=self.a.mat, b=self.L.vec, pre=self.pre, freedofs=self.fes.FreeDofs(),
                              x=self.gfu.vec, tol=self.linear_tolerance, maxsteps=self.linear_max_iterations,
                              restart=100)
        else:
            raise ValueError(self.linear_solver)
        self.linear_iterations.append(getattr(be, 'last_iters', 0))

    def step(self) -> float:
        """One time step of the transient solve; returns the last Picard update norm."""
        self.t.Set(self.t.Get() + self.dt.Get())
        u_comp = self.gfu.components[0]
        it = 1
        err = 0.0
        while True:
            self.apply_dirichlet_bcs()
            self.assemble()
            self.linear_solve()
            diff = self.W - u_comp
            err = float(np.sqrt(self._integrate(ngs.InnerProduct(diff, diff))))
            unorm = float(np.sqrt(self._integrate(ngs.InnerProduct(u_comp, u_comp))))
            it += 1
            self.W.vec.data = u_comp.vec
            if err < self.abs_nonlinear_tolerance + self.rel_nonlinear_tolerance * unorm or \
                    it > self.nonlinear_max_iters:
                break
        self.picard_iterations = it - 1
        self.gfu_0.vec.data = self.gfu.vec
        return err

    def _integrate(self, cf):
        return ngs.Integrate(cf, self.mesh)

    def errors(self):
        """L2 errors against the reference solution (helpers/error.py:54-128); pressure compared up to its mean."""
        u, p = self.gfu.components
        du = u - self.u_ref
        eu = float(np.sqrt(self._integrate(ngs.InnerProduct(du, du))))
        area = self._integrate(ngs.CoefficientFunction(1.0))
        pm = self._integrate(p - self.p_ref) / area
        dp = p - self.p_ref - pm
        ep = float(np.sqrt(self._integrate(dp * dp)))
        return eu, ep
