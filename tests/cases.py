"""Shared weak-form builders for the parity tests.

Each builder writes the forms the way the reference's model classes do (file:line cited per builder) against the
NGSolve-style front end, so the same code runs on the CUDA backend and on the oracle.
"""
import numpy as np

from opencmp_b200.mesh import delaunay_rectangle, structured_2d


def _ngs():
    import opencmp_b200.ngs as ngs
    return ngs


def dg_funcs(ngs, mesh, nu):
    """reference helpers/ngsolve_.py:48-70"""
    n = ngs.specialcf.normal(mesh.dim)
    h = ngs.specialcf.mesh_size
    return n, h, nu / h


def jump(q):
    return q - q.Other()


def grad_avg(ngs, q):
    return 0.5 * (ngs.Grad(q) + ngs.Grad(q.Other()))


def avg(q):
    return 0.5 * (q + q.Other())


def poisson(mesh, order, DG, family='H1', transient_dt=None):
    """Poisson forms: reference opencmp/models/poisson.py:71-158 (+ mass term of base_model.py:418-497)."""
    ngs = _ngs()
    m = ngs.Mesh(mesh)
    dnames = 'top|bottom'
    fes = ngs.FESpace([getattr(ngs, family)(m, order=order, dirichlet=dnames, dgjumps=DG)], dgjumps=DG)
    u, v = fes.TrialFunction()[0], fes.TestFunction()[0]
    dt = ngs.Parameter(1.0 if transient_dt is None else transient_dt)
    t = ngs.Parameter(0.25)
    dc = ngs.CoefficientFunction(0.5)
    exact = ngs.sin(ngs.pi * ngs.x) * ngs.cos(ngs.pi * ngs.y) * ngs.exp(-t)
    f = 2 * ngs.pi ** 2 * dc * exact
    gradex = ngs.CoefficientFunction((ngs.pi * ngs.cos(ngs.pi * ngs.x) * ngs.cos(ngs.pi * ngs.y) * ngs.exp(-t),
                                      -ngs.pi * ngs.sin(ngs.pi * ngs.x) * ngs.sin(ngs.pi * ngs.y) * ngs.exp(-t)))
    n, h, alpha = dg_funcs(ngs, m, 10.0 * order ** 2)
    a = ngs.BilinearForm(fes)
    a += dt * dc * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) * ngs.dx
    L = ngs.LinearForm(fes)
    L += dt * f * v * ngs.dx
    ds_d = ngs.ds(skeleton=DG, definedon=m.Boundaries(dnames))
    ds_n = ngs.ds(skeleton=DG, definedon=m.Boundaries('left|right'))
    if DG:
        a += dt * dc * (-u * n * ngs.Grad(v) - ngs.Grad(u) * n * v + alpha * u * v) * ds_d
        ju, jv = jump(u), jump(v)
        a += dt * dc * (-ju * n * grad_avg(ngs, v) - grad_avg(ngs, u) * n * jv + alpha * ju * jv) * ngs.dx(skeleton=True)
        L += dt * dc * (alpha * exact * v - exact * n * ngs.Grad(v)) * ds_d
    L += dt * dc * (gradex * n) * v * ds_n          # Neumann data
    if transient_dt is not None:
        a += u * v * ngs.dx
    gfu = ngs.GridFunction(fes)
    return dict(ngs=ngs, mesh=m, fes=fes, a=a, L=L, gfu=gfu, exact=exact, dnames=dnames, params=(dt, t))


def stokes(mesh, order, DG, wind=None, dt_val=1.0, mass=False, walls='wall', nu=1e-3, RT=False):
    """Stokes / Oseen forms: reference opencmp/models/stokes.py:43-129 and models/ins.py:178-321."""
    ngs = _ngs()
    m = ngs.Mesh(mesh)
    if DG:
        V = ngs.HDiv(m, order=order, dirichlet=walls, dgjumps=True, RT=RT)        # models/ins.py:114-117
        Q = ngs.L2(m, order=order - (0 if RT else 1), dgjumps=True)
    else:
        V = ngs.VectorH1(m, order=order, dirichlet=walls)
        Q = ngs.H1(m, order=order - 1)
    fes = ngs.FESpace([V, Q], dgjumps=DG)
    (u, p), (v, q) = fes.TrialFunction(), fes.TestFunction()
    dt = ngs.Parameter(dt_val)
    kv = ngs.CoefficientFunction(nu)
    n, h, alpha = dg_funcs(ngs, m, 10.0 * order ** 2)
    uex = ngs.CoefficientFunction((5.0 / (2 * nu) * (0.2 * ngs.y - ngs.y * ngs.y), 0.0))
    pex = 5.0 * (1.0 - ngs.x) + 5.0
    a = ngs.BilinearForm(fes)
    a += dt * kv * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) * ngs.dx
    a += dt * (-ngs.div(u) * q - ngs.div(v) * p - 1e-10 * p * q) * ngs.dx
    L = ngs.LinearForm(fes)
    L += dt * v * ngs.CoefficientFunction((0.0, 0.0)) * ngs.dx
    W = None
    if wind is not None:
        W = ngs.GridFunction(V)
        W.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(wind(V.ndof)))
        a += -dt * ngs.InnerProduct(ngs.OuterProduct(u, W), ngs.Grad(v)) * ngs.dx
    ds_w = ngs.ds(skeleton=DG, definedon=m.Boundaries(walls))
    if DG:
        a += dt * (kv * alpha * u * v - kv * ngs.InnerProduct(ngs.Grad(u), ngs.OuterProduct(v, n))
                   - kv * ngs.InnerProduct(ngs.Grad(v), ngs.OuterProduct(u, n))) * ds_w
        ju, jv = jump(u), jump(v)
        a += dt * (kv * alpha * ngs.InnerProduct(ju, jv)
                   - kv * ngs.InnerProduct(grad_avg(ngs, u), ngs.OuterProduct(jv, n))
                   - kv * ngs.InnerProduct(grad_avg(ngs, v), ngs.OuterProduct(ju, n))) * ngs.dx(skeleton=True)
        L += dt * (kv * alpha * uex * v - kv * ngs.InnerProduct(ngs.Grad(v), ngs.OuterProduct(uex, n))) * ds_w
        if W is not None:
            a += dt * v * (0.5 * W * n * u + 0.5 * ngs.Norm(W * n) * u) * ds_w
            a += dt * jv * (W * n * avg(u) + 0.5 * ngs.Norm(W * n) * ju) * ngs.dx(skeleton=True)
            L += dt * v * (-0.5 * W * n * uex + 0.5 * ngs.Norm(W * n) * uex) * ds_w
    for marker, pval in (('inlet', 10.0), ('outlet', 5.0)):
        dsm = ngs.ds(skeleton=DG, definedon=m.Boundaries(marker))
        L += dt * v * (-pval * n) * dsm
        if W is not None:
            a += dt * v * (ngs.IfPos(W * n, W * n, 0.0) * u) * dsm
    if mass:
        a += u * v * ngs.dx
    gfu = ngs.GridFunction(fes)
    return dict(ngs=ngs, mesh=m, fes=fes, a=a, L=L, gfu=gfu, uex=uex, pex=pex, walls=walls, W=W, V=V, params=(dt,))


def channel_mesh(n=12, seed=3):
    return delaunay_rectangle(n, seed, size=(1.0, 0.2), names=('wall', 'outlet', 'wall', 'inlet'))


def square_mesh(n=6, seed=1):
    return delaunay_rectangle(n, seed)


def random_wind(ndof, seed=5, scale=0.3):
    return scale * np.random.default_rng(seed).uniform(-1.0, 1.0, ndof)


def direct_solve(case, be_numpy=True):
    """``Model.linear_solve`` direct branch, reference base_model.py:918-922."""
    ngs = case['ngs']
    a, L, gfu, fes = case['a'], case['L'], case['gfu'], case['fes']
    inv = a.mat.Inverse(fes.FreeDofs())
    r = L.vec.CreateVector()
    r.data = L.vec - a.mat * gfu.vec
    gfu.vec.data += inv * r
    return gfu.vec.NumPy().copy()


def poisson_dim(mesh, order, DG=False):
    """Diffuse-interface Poisson: reference opencmp/models/poisson_dim.py:33-160 — every integrand weighted by the
    phase field phi, Nitsche terms on the diffuse boundary as volume terms with grad(phi), |grad(phi)| and a mask."""
    ngs = _ngs()
    m = ngs.Mesh(mesh)
    fes = ngs.FESpace([ngs.H1(m, order=order, dgjumps=DG)], dgjumps=DG)
    H = ngs.H1(m, order=order)
    u, v = fes.TrialFunction()[0], fes.TestFunction()[0]
    dt = ngs.Parameter(1.0)
    dc = ngs.CoefficientFunction(0.7)
    r2 = (ngs.x - 0.5) * (ngs.x - 0.5) + (ngs.y - 0.5) * (ngs.y - 0.5)
    lam = 0.08
    phi_cf = 0.5 * (1.0 + ngs.erf((0.3 - ngs.sqrt(r2 + 1e-12)) / lam))
    phi, mag, mask = ngs.GridFunction(H), ngs.GridFunction(H), ngs.GridFunction(H)
    phi.Set(phi_cf)
    mask.Set(ngs.CoefficientFunction(1.0))
    gphi = ngs.Grad(phi)
    mag.Set(ngs.sqrt(gphi * gphi + 1e-14))
    n, h, alpha = dg_funcs(ngs, m, 10.0 * order ** 2)
    g = ngs.sin(ngs.x) + ngs.y
    f = ngs.CoefficientFunction(1.0) + ngs.x * ngs.y
    a = ngs.BilinearForm(fes)
    a += dt * dc * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) * phi * ngs.dx
    a += dt * alpha * u * v * (1.0 - phi) * ngs.dx                                  # poisson_dim.py:47
    a += dt * dc * (ngs.Grad(u) * gphi * v + ngs.Grad(v) * gphi * u + alpha * u * v * mag) * mask * ngs.dx
    if DG:
        ju, jv = jump(u), jump(v)
        a += dt * dc * (-ju * n * grad_avg(ngs, v) - grad_avg(ngs, u) * jv * n + alpha * ju * jv) * phi \
            * ngs.dx(skeleton=True)
    L = ngs.LinearForm(fes)
    L += dt * f * v * phi * ngs.dx
    L += dt * dc * (ngs.Grad(v) * gphi * g + alpha * g * v * mag) * mask * ngs.dx   # poisson_dim.py:130-136
    gfu = ngs.GridFunction(fes)
    return dict(ngs=ngs, mesh=m, fes=fes, a=a, L=L, gfu=gfu, noset=True, keep=(phi, mag, mask))


def species(mesh, order, wind):
    """Multi-component species transport, DG: reference opencmp/models/multi_component_ins.py:158-449 — diffusion SIP,
    upwinded advection by a velocity field (IfPos max/min), cross-species linear reactions, outflow boundary terms."""
    ngs = _ngs()
    m = ngs.Mesh(mesh)
    V = ngs.HDiv(m, order=order, dgjumps=True)
    A, B = ngs.L2(m, order=order, dgjumps=True), ngs.L2(m, order=order, dgjumps=True)
    fes = ngs.FESpace([A, B], dgjumps=True)
    (ca, cb), (ra, rb) = fes.TrialFunction(), fes.TestFunction()
    w = ngs.GridFunction(V)
    w.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(wind(V.ndof)))
    dt = ngs.Parameter(0.01)
    n, h, alpha = dg_funcs(ngs, m, 10.0 * order ** 2)
    a = ngs.BilinearForm(fes)
    L = ngs.LinearForm(fes)
    for c, r, D, k, src in ((ca, ra, 0.1, 0.5, 1.0 + ngs.x), (cb, rb, 0.02, -0.5, ngs.sin(ngs.y))):
        a += c * r * ngs.dx                                                              # time derivative
        a += dt * D * ngs.InnerProduct(ngs.Grad(c), ngs.Grad(r)) * ngs.dx              # :223-224
        a += -dt * c * (w * ngs.Grad(r)) * ngs.dx                                       # :260
        jc, jr = jump(c), jump(r)
        a += dt * D * (alpha * jc * jr - grad_avg(ngs, c) * n * jr - grad_avg(ngs, r) * n * jc) * ngs.dx(skeleton=True)
        wn = w * n
        a += dt * jr * (c * ngs.IfPos(wn, wn, 0.0) + c.Other() * ngs.IfPos(wn, 0.0, wn)) * ngs.dx(skeleton=True)
        a += dt * r * c * ngs.IfPos(wn, wn, 0.0) * ngs.ds(skeleton=True)                # :267-275
        L += dt * src * r * ngs.dx
    a += -dt * 0.5 * ca * rb * ngs.dx + dt * 0.25 * cb * ra * ngs.dx                    # cross-species reactions :245-256
    gfu = ngs.GridFunction(fes)
    return dict(ngs=ngs, mesh=m, fes=fes, a=a, L=L, gfu=gfu, noset=True, keep=(w,))


def stokes_3d(cell, order, n=2):
    """Taylor-Hood Stokes on a structured 3-D box (hex Q2/Q1 is the SURVEY 8(d) 3-D throughput configuration)."""
    from opencmp_b200.mesh import structured_3d
    ngs = _ngs()
    m = ngs.Mesh(structured_3d([n, n, n], cell=cell))
    V = ngs.VectorH1(m, order=order, dirichlet='back|left|bottom')
    Q = ngs.H1(m, order=order - 1)
    fes = ngs.FESpace([V, Q])
    (u, p), (v, q) = fes.TrialFunction(), fes.TestFunction()
    W = ngs.GridFunction(V)
    W.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(random_wind(V.ndof, 9)))
    a = ngs.BilinearForm(fes)
    a += (0.1 * ngs.InnerProduct(ngs.Grad(u), ngs.Grad(v)) - ngs.div(u) * q - ngs.div(v) * p - 1e-10 * p * q) * ngs.dx
    a += -ngs.InnerProduct(ngs.OuterProduct(u, W), ngs.Grad(v)) * ngs.dx
    n3 = ngs.specialcf.normal(3)
    a += v * (ngs.IfPos(W * n3, W * n3, 0.0) * u) * ngs.ds(definedon=m.Boundaries('front|top'))
    L = ngs.LinearForm(fes)
    L += v * ngs.CoefficientFunction((ngs.z, ngs.sin(ngs.x), ngs.y * ngs.y)) * ngs.dx
    L += v * (-2.0 * n3) * ngs.ds(definedon=m.Boundaries('right'))
    gfu = ngs.GridFunction(fes)
    return dict(ngs=ngs, mesh=m, fes=fes, a=a, L=L, gfu=gfu, noset=True, keep=(W,))


def ins_dim_3d(n=4, n0=2, **kw):
    """BASELINE config 5 at test size: INS with the diffuse-interface method (reference models/ins_dim.py) on the
    structured hex box, Taylor-Hood Q2/Q1, phase field of a sphere — the workload bench.py --workload ins3d_dim runs."""
    from opencmp_b200.mesh import structured_3d
    from opencmp_b200.workloads import INSSphereDIM3D
    ngs = _ngs()
    mesh = structured_3d([n0] * 3, scale=(2.0,) * 3, offset=(1.0,) * 3)
    k = n0
    while k < n:
        mesh.Refine()
        k *= 2
    w = INSSphereDIM3D(n, mesh=mesh, **kw)
    w.W.vec.data = ngs.BaseVector(ngs.get_backend().from_numpy(
        ngs.get_backend().to_numpy(w.W.vec.a) + random_wind(w.V.ndof, 13, 0.1) * w.V.FreeDofs()))
    return dict(ngs=ngs, mesh=w.mesh, fes=w.fes, a=w.a, L=w.L, gfu=w.gfu, noset=True, keep=(w,), workload=w)


def quad_channel_mesh(nx=8, ny=4):
    """1 x 0.2 channel of quadrilaterals with the boundary names of pytests/mesh_files/channel_3bcs.vol."""
    from opencmp_b200.mesh import structured_2d
    m = structured_2d([nx, ny], scale=(1.0, 0.2), cell='quad')
    m.bnd_names = [{'bottom': 'wall', 'top': 'wall', 'left': 'inlet', 'right': 'outlet'}[n] for n in m.bnd_names]
    return m
