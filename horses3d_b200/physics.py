"""H3dPhysics (include/h3d_gpu.h) and its construction from control-file values.

Follows ConstructPhysicsStorage_NS (libs/physics/navierstokes/PhysicsStorage_NS.f90:143-435) and
SetRiemannSolver (RiemannSolvers_NS.f90:120-286): same defaults, same arithmetic for the derived constants.
"""
import ctypes as C

STANDARD_DG, SPLIT_DG = 0, 1
RIEMANN = {"roe": 0, "lax-friedrichs": 1, "central": 2, "rusanov": 3, "standard roe": 4, "u-diss": 5, "roe-pike": 6,
           "low dissipation roe": 7, "matrix dissipation": 8}
AVERAGING = {"standard": 0, "kennedy-gruber": 1, "pirozzoli": 2, "ducros": 3, "morinishi": 4, "entropy conserving": 5,
             "chandrasekar": 6}
LES = {"none": 0, "smagorinsky": 1, "wale": 2, "vreman": 3}
VISCOUS = {"br1": 0, "br2": 1, "ip": 2}
IP_VARIANT = {"sipg": -1, "iipg": 0, "nipg": 1}
GRADVARS = {"state": 0, "entropy": 1, "energy": 2}

INT_VOLUME, INT_KINETIC_ENERGY, INT_KINETIC_ENERGY_RATE, INT_ENSTROPHY = 0, 1, 2, 3
INT_VELOCITY, INT_ENTROPY, INT_ENTROPY_RATE, INT_INTERNAL_ENERGY, INT_ENTROPY_BALANCE, INT_MATH_ENTROPY = 4, 5, 6, 7, 8, 9
INT_KINETIC_ENERGY_BALANCE = 10
RK3, RK5 = 3, 5
EULER, LSERK14_4, SSPRK33, SSPRK43 = 1, 14, 33, 43
SURF_SURFACE, SURF_MASS_FLOW, SURF_FLOW_RATE, SURF_PRESSURE, SURF_VEC_SURFACE, SURF_TOTAL_FORCE, SURF_PRESSURE_FORCE, SURF_VISCOUS_FORCE = range(8)
FACE_INTERIOR, FACE_BOUNDARY, FACE_MPI = 1, 2, 3
BC_TYPES = {"periodic": 0, "noslipwall": 1, "freeslipwall": 2, "inflow": 3, "outflow": 4}


class H3dPhysics(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "gamma", "gammaMinus1", "Mach", "Re", "Pr", "mu", "kappa", "mu_to_kappa", "gammaM2",
        "S_div_Tref", "T_renorm", "lambdaStab", "smagorinsky_Cs", "Prt", "penaltyParameter")] + [(k, C.c_int) for k in (
        "flowIsNavierStokes", "computeGradients", "inviscid", "riemann", "averaging", "les", "les_wall_model", "viscous", "ipVariant", "gradientVariables")]


def make_physics(flow="NS", mach=0.08, reynolds=1600.0, prandtl=0.72, inviscid="standard", riemann="roe",
                 averaging="standard", lambda_stab=1.0, compute_gradients=None, les="none", smagorinsky_cs=None, les_wall_model="none",
                 sutherland_temperature=None, reference_temperature=None, sutherland_ref_temperature=None,
                 viscous="BR1", penalty_parameter=None, ip_variant="SIPG", gradient_variables="State"):
    p = H3dPhysics()
    gamma = 1.4
    gm1 = 1.4 - 1.0                                   # PhysicsStorage_NS.f90:120 (not 0.4)
    p.gamma, p.gammaMinus1, p.Mach, p.Pr, p.Prt = gamma, gm1, mach, prandtl, prandtl
    ns = flow.lower() != "euler"
    if ns:
        p.Re = reynolds
        if reynolds != 0.0:
            p.mu = 1.0 / reynolds                      # :240
            p.kappa = 1.0 / (gm1 * (mach * mach) * reynolds * prandtl)   # :241-243
        p.mu_to_kappa = 1.0 / (gm1 * (mach * mach) * prandtl)            # :250
    if les.lower() != "none":
        ns = True                                      # :405-414
        compute_gradients = True
    p.flowIsNavierStokes = int(ns)
    if compute_gradients is None:
        compute_gradients = ns                         # :255-275
    p.computeGradients = int(bool(compute_gradients) or ns)
    p.gammaM2 = gamma * (mach * mach)                  # :285
    Tref = reference_temperature if reference_temperature is not None else 520.0 * 5.0 / 9.0       # :294
    S = sutherland_temperature if sutherland_temperature is not None else 198.6 * 5.0 / 9.0          # :419
    TrefS = sutherland_ref_temperature if sutherland_ref_temperature is not None else Tref         # :425
    p.S_div_Tref = S / TrefS
    p.T_renorm = Tref / TrefS
    p.inviscid = {"standard": STANDARD_DG, "split-form": SPLIT_DG}[inviscid.lower()]
    p.riemann = RIEMANN[riemann.lower()]
    p.averaging = AVERAGING[averaging.lower()]
    p.lambdaStab = 0.0 if p.riemann == RIEMANN["central"] else lambda_stab    # RiemannSolvers_NS.f90:211-224
    p.viscous = VISCOUS[viscous.lower()]
    # "penalty parameter": BR2 eta defaults to 2 (EllipticBR2.f90:80-91), IP sigma to 1 (EllipticIP.f90:110-121)
    p.penaltyParameter = penalty_parameter if penalty_parameter is not None else {0: 0.0, 1: 2.0, 2: 1.0}[p.viscous]
    p.ipVariant = IP_VARIANT[ip_variant.lower()]
    p.gradientVariables = GRADVARS[gradient_variables.lower()] if ns else 0       # SpatialDiscretization.f90:106-148, 190-193
    p.les = LES[les.lower()]
    # "LES model intensity" defaults: Smagorinsky 0.2, WALE 0.325, Vreman 0.07 (LESModels.f90:233-254, 337-356, 466-485)
    p.smagorinsky_Cs = smagorinsky_cs if smagorinsky_cs is not None else {0: 0.2, 1: 0.2, 2: 0.325, 3: 0.07}[p.les]
    p.les_wall_model = {"none": 0, "linear": 1}[les_wall_model.lower()]      # LESModels.f90:137-165
    return p


def bc_parameters(kind, physics, **kw):
    """16 parameters of one boundary zone, laid out as include/h3d_gpu.h documents (walls, inflow, outflow).

    Mirrors the BC constructors: NoSlipWallBC.f90:150-210, FreeSlipWallBC.f90:150-200, InflowBC.f90:150-330,
    OutflowBC.f90:120-200.  Non-dimensional inputs (the reference divides by refValues there)."""
    import math
    P = [0.0] * 16
    kind = kind.lower()
    Tref = 520.0 * 5.0 / 9.0
    if kind in ("noslipwall", "freeslipwall"):
        vw = kw.get("wall_velocity", (0.0, 0.0, 0.0))
        if kind == "noslipwall":
            P[0:3] = [float(x) for x in vw]
        iso = 0.0 if kw.get("adiabatic", True) else 1.0
        Twall = kw.get("wall_temperature", Tref) if iso else 0.0
        P[3], P[4] = iso, Twall
        P[5] = Tref * physics.gammaM2 * physics.gammaMinus1 if kind == "noslipwall" else Tref * physics.gammaM2
        P[6] = Twall / (Tref * physics.gammaMinus1 * physics.gammaM2) if iso else 0.0
    elif kind == "inflow":
        # defaults of InflowBC: rho = 1, |v| = 1, p = 1/(gamma M^2); AoA from the control file
        rho = kw.get("rho", 1.0)
        vmag = kw.get("v", 1.0)
        th, ph = kw.get("aoa_theta", 0.0), kw.get("aoa_phi", 0.0)
        u = vmag * math.cos(th) * math.cos(ph)
        v = vmag * math.sin(th) * math.cos(ph)
        w = vmag * math.sin(ph)
        P[0], P[1], P[2], P[3] = rho, u, v, w
        P[4] = kw.get("p", 1.0 / physics.gammaM2)
    elif kind == "outflow":
        P[4] = kw.get("p", 1.0 / physics.gammaM2)
    elif kind == "periodic":
        pass
    else:
        raise ValueError("boundary condition %r is not implemented" % kind)
    return P
