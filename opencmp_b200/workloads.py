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
        self.pre = ngs.Preconditioner(self.a, preconditioner) if preconditioner is not None else None
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


def ins_dim_cg_forms(fes, phi, gphi, mag, mask, w, gfu_0, g, f, kv, alpha, dtp):
    """Bilinear and linear form of one implicit-Euler / Oseen step of the reference's ``INSDIM`` model on a conforming
    (Taylor-Hood) space, no rigid-body motion, one DIM Dirichlet marker — a restatement of

      construct_bilinear_time_coefficient   opencmp/models/ins_dim.py:106-146
      construct_bilinear_time_ODE           opencmp/models/ins_dim.py:33-104
      construct_linear                      opencmp/models/ins_dim.py:148-206
      time-derivative terms x phi           solvers/time_integration_schemes.py:663-685, models/base_model.py:418-497

    checked entry by entry against the matrix / right-hand side the reference class itself assembles
    (tests/test_reference_models.py::test_workload_forms_equal_reference_insdim). ``gphi`` / ``mag`` are the DIM
    solver's grad(phi) and |grad(phi)|: expressions of phi when it was generated on the simulation mesh
    (diffuse_interface/dim.py:391-392), separately projected GridFunctions otherwise (:393-409)."""
    (u, p), (v, q) = fes.TrialFunction(), fes.TestFunction()
    a = ngs.BilinearForm(fes)
    a += dtp * (-ngs.div(u) * q - ngs.div(v) * p - 1e-10 * p * q) * phi * ngs.dx
    a += -dtp * p * ngs.div(v) * (1.0 - phi) * ngs.dx
    a += dtp * (kv * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v))) * phi * ngs.dx
    a += -dtp * ngs.InnerProduct(ngs.OuterProduct(u, w), ngs.Grad(v)) * phi * ngs.dx
    a += dtp * alpha * u * v * (1.0 - phi) * ngs.dx
    a += dtp * (kv * ngs.InnerProduct(ngs.Grad(u), ngs.OuterProduct(v, gphi))
                + kv * ngs.InnerProduct(ngs.Grad(v), ngs.OuterProduct(u, gphi))
                + kv * alpha * u * v * mag) * mask * ngs.dx
    a += dtp * (v * (0.5 * w * (-gphi) * u + 0.5 * ngs.Norm(w * (-gphi)) * u)) * mask * ngs.dx
    a += (u * v) * phi * ngs.dx
    L = ngs.LinearForm(fes)
    L += dtp * v * f * phi * ngs.dx
    L += dtp * (kv * ngs.InnerProduct(ngs.Grad(v), ngs.OuterProduct(g, gphi))
                + kv * alpha * g * v * mag) * mask * ngs.dx
    L += dtp * v * (0.5 * w * gphi * g + 0.5 * ngs.Norm(w * (-gphi)) * g) * mask * ngs.dx
    L += (gfu_0.components[0] * v) * phi * ngs.dx
    return a, L


class INSSphereDIM3D:
    """3-D incompressible Navier-Stokes with the diffuse-interface method (BASELINE config 5).

    Restates what the reference executes per time step for its ``INSDIM`` model with Oseen linearisation and the
    implicit-Euler scheme, conforming (CG) Taylor-Hood Q2/Q1 on the structured hexahedral box [-1,1]^3 of the DIM mesh
    generator (``quad_mesh = True``):

      forms      opencmp/models/ins_dim.py:33-196 (phi-weighted stress / convection / incompressibility, penalisation
                 alpha u v (1-phi) and -p div v (1-phi), Nitsche terms on the diffuse boundary with grad(phi),
                 |grad(phi)| and the mask, their right-hand-side analogues)
      DIM fields opencmp/diffuse_interface/dim.py:390-445 (phi as an H1 GridFunction of the interpolant order clamped
                 to [1e-10, 1]; grad(phi) and |grad(phi)| evaluated from it; mask = 1)
      phi        erf profile of the signed distance, opencmp/diffuse_interface/interface.py:173-178, here of a sphere
                 of radius R computed analytically instead of through the STL -> EDT pipeline (out of scope)
      scheme / time step / Picard / solve: as INSTaylorGreen above.

    Physical set-up: fluid inside the sphere, whose (diffuse) wall rotates rigidly about the z axis: DIM Dirichlet
    data g = omega e_z x r. Rigid rotation is the exact Navier-Stokes solution in the sphere, which the workload starts
    from, so ``errors()`` measures how well the DIM step preserves it. Homogeneous conformal Dirichlet data on the box.
    """

    def __init__(self, N: int, order: int = 2, dt: float = 1e-2, nu: float = 1.0, ipc: float = 10.0,
                 radius: float = 0.5, lam: float = 0.25, lam_cells: float = None, omega_rot: float = 1.0,
                 preconditioner: str = 'multigrid', linear_tolerance: float = 1e-12, linear_max_iterations: int = 400,
                 nonlinear_max_iterations: int = 3, nonlinear_tolerance=(1e-4, 1e-6), n0: int = 2, mesh=None,
                 integrate=None, periodic=(True, False, False), wall_period: float = None, wall_amp: float = 0.5):
        from .mesh import structured_3d
        if integrate is not None:              # element-partitioned runs: owned cells + all-reduce
            self._integrate = integrate
        if mesh is None:
            nc = N
            while preconditioner == 'multigrid' and nc % 2 == 0 and nc > n0:
                nc //= 2
            mesh = structured_3d([nc] * 3, scale=(2.0,) * 3, offset=(1.0,) * 3)
            while nc < N:
                mesh.Refine()
                nc *= 2
        self.mesh = ngs.Mesh(mesh)
        m = self.mesh
        k = order
        self.order = k
        self.preconditioner = preconditioner
        self.linear_tolerance, self.linear_max_iterations = linear_tolerance, linear_max_iterations
        self.nonlinear_max_iters = nonlinear_max_iterations
        self.rel_nonlinear_tolerance, self.abs_nonlinear_tolerance = nonlinear_tolerance
        dnames = 'back|left|front|right|bottom|top'
        self.dirichlet = dnames
        # models/ins.py:95-128 with DG = False: Taylor-Hood
        self.V = ngs.VectorH1(m, order=k, dirichlet=dnames)
        self.Q = ngs.H1(m, order=k - 1)
        self.fes = ngs.FESpace([self.V, self.Q])
        self.t, self.dt = ngs.Parameter(0.0), ngs.Parameter(dt)
        dtp = self.dt
        x, y, z = ngs.x, ngs.y, ngs.z
        kv = ngs.CoefficientFunction(nu)
        h = ngs.specialcf.mesh_size
        alpha = (ipc * k ** 2) / h                          # base_model.py:150, helpers/ngsolve_.py:67
        # ---- diffuse-interface fields (dim.py:390-445) ----
        # interface width: a physical length (default R/2), or ``lam_cells`` mesh widths. The reference clamps phi at
        # 1e-10, which leaves the pressure outside the fluid determined only through 1e-10-scaled rows; once a large
        # part of the box is clamped (lam <~ 0.15 here) the discrete system is numerically singular — two refinement
        # passes of a sparse LU disagree by O(1) — so no solver, direct or Krylov, has a meaningful answer there.
        self.lam = lam if lam_cells is None else lam_cells * 2.0 / N
        H = ngs.H1(m, order=k)
        self.fes_phi = H
        # one sphere per [-1,1]^3 brick: the element-partitioned run tiles the domain with bricks along the
        # ``periodic`` directions (weak scaling); local coordinates = position inside the brick
        wrap = lambda c: (c + 1.0) - 2.0 * ngs.floor(0.5 * (c + 1.0)) - 1.0
        xl = wrap(x) if periodic[0] else x
        yl = wrap(y) if periodic[1] else y
        zl = wrap(z) if periodic[2] else z
        r = ngs.sqrt(xl * xl + yl * yl + zl * zl + 1e-30)
        phi_cf = 0.5 * (1.0 + ngs.erf((radius - r) / self.lam))
        self.phi = ngs.GridFunction(H)
        self.phi.Set(phi_cf)
        be = ngs.get_backend()
        pv = be.to_numpy(self.phi.vec.a)
        self.phi.vec.data = ngs.BaseVector(be.from_numpy(np.clip(pv, 1e-10, 1.0)))      # dim.py:434-435
        self.mask = ngs.GridFunction(H)
        self.mask.Set(ngs.CoefficientFunction(1.0))
        phi, mask = self.phi, self.mask
        # wall speed: constant (default), or oscillating omega(t) = omega_rot (1 + wall_amp sin(2 pi t / wall_period)).
        # With a constant wall the flow relaxes to the rigid rotation and, a few steps in, the previous solution is such
        # a good initial guess that a reduction of the preconditioned residual by 1e-12 RELATIVE TO THE INITIAL ONE
        # (the stopping rule of ngsolve.solvers.GMRes) falls below the round-off floor of the residual evaluation: the
        # solves then run to maxsteps. A time-dependent wall keeps every step equally hard, which is what a
        # throughput measurement over many steps needs (bench.py uses wall_period = 10 dt).
        if wall_period:
            omega_t = omega_rot * (1.0 + wall_amp * ngs.sin((2.0 * np.pi / wall_period) * self.t))
        else:
            omega_t = ngs.CoefficientFunction(omega_rot)
        self.u_ref = ngs.CoefficientFunction((-omega_t * yl, omega_t * xl, 0.0 * x))
        self.p_ref = 0.5 * omega_t * omega_t * (xl * xl + yl * yl)
        g = self.u_ref
        f = ngs.CoefficientFunction((0.0, 0.0, 0.0))
        self.gfu, self.gfu_0 = ngs.GridFunction(self.fes), ngs.GridFunction(self.fes)
        self.W = ngs.GridFunction(self.V)                    # Oseen wind, models/ins.py:130-134
        w = self.W
        gphi = ngs.Grad(phi)                                 # dim.py:391-392 (phi lives on the simulation mesh)
        self.a, self.L = ins_dim_cg_forms(self.fes, phi, gphi, ngs.Norm(gphi), mask, w, self.gfu_0, g, f, kv, alpha,
                                          dtp)
        self.pre = ngs.Preconditioner(self.a, preconditioner) if preconditioner is not None else None
        # start from the rigid rotation inside the sphere (zero on the box boundary)
        cut = 0.5 * (1.0 + ngs.erf((0.9 - r) / 0.05))
        self.gfu.components[0].Set(self.u_ref * cut)
        self.gfu.components[1].Set(self.p_ref * cut)
        self.apply_dirichlet_bcs()
        self.gfu_0.vec.data = self.gfu.vec
        self.W.vec.data = self.gfu.components[0].vec
        self.picard_iterations = 0
        self.linear_iterations = []

    # GMRES restart length. The solves of this workload need 40-60 iterations on one brick but 100+ on a row of
    # bricks (one nearly undetermined pressure mode per sphere); a restart in the middle of a solve throws the Krylov
    # space away and costs 15-25 % more iterations (measured on the CPU restatement, profiles/r1_solver_convergence.md),
    # so the basis is allowed to grow to 200 vectors (4.6 GB at 2.9 M DOFs per GPU).
    gmres_restart = 200

    ndof = INSTaylorGreen.ndof
    nnz = INSTaylorGreen.nnz
    assemble = INSTaylorGreen.assemble
    step = INSTaylorGreen.step
    _integrate = INSTaylorGreen._integrate

    def apply_dirichlet_bcs(self):
        """base_model.py:321-341 — homogeneous data on the box faces. The coefficient function is the SAME object on
        every call, like the entries of the reference's BC dictionary: the lowered projection (program, launch plans,
        contributor lists) is cached on it — a fresh object per call cost 72 ms per projection at 48^3."""
        if getattr(self, '_g_box', None) is None:
            self._g_box = ngs.CoefficientFunction((0.0, 0.0, 0.0))
            self._box_region = self.mesh.Boundaries(self.dirichlet)
        self.gfu.components[0].Set(self._g_box, definedon=self._box_region)

    def linear_solve(self):
        """base_model.py:886-947, linear_solver = GMRes"""
        be = ngs.get_backend()
        ngs.solvers.GMRes(A=self.a.mat, b=self.L.vec, pre=self.pre, freedofs=self.fes.FreeDofs(), x=self.gfu.vec,
                          tol=self.linear_tolerance, maxsteps=self.linear_max_iterations, restart=self.gmres_restart)
        self.linear_iterations.append(getattr(be, 'last_iters', 0))

    def errors(self):
        """phi-weighted L2 deviation from the rigid rotation, relative to its phi-weighted norm."""
        u = self.gfu.components[0]
        du = u - self.u_ref
        eu = float(np.sqrt(self._integrate(ngs.InnerProduct(du, du) * self.phi)))
        un = float(np.sqrt(self._integrate(ngs.InnerProduct(self.u_ref, self.u_ref) * self.phi)))
        return eu / un, un
