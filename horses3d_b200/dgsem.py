"""Host-side mirror of the reference's driver objects for the explicit NS path.

`DGSem` plays the part of the reference's DGSem + TimeIntegrator_t for the hot path only: it hands mesh,
basis and physics to a backend that implements the C ABI of include/h3d_gpu.h and exposes the reference's
procedure names with the reference's argument meaning:

    ComputeTimeDerivative(time)              SpatialDiscretization.f90:227   (ComputeTimeDerivative_f)
    TakeRK3Step(t, dt) / TakeRK5Step(t, dt)  ExplicitMethods.f90:667, 790    (TimeStep_FCN)
    MaxTimeStep(cfl, dcfl)                   DGSEMClass.f90:870
    ComputeMaxResiduals()                    DGSEMClass.f90:770
    ScalarVolumeIntegral(kind)               VolumeIntegrals.f90:76
    checkForNan()                            ExplicitMethods.f90:1856
    integrate(...)                           TimeIntegrator.f90:667-673, 737-959 (explicit branch)

The backend is an `Api` (capi.py).  The product backend is GpuApi; tests pass the CPU oracle's Api to run
the very same driver code against the restated reference algorithm.
"""
import ctypes as C

import numpy as np

from . import physics as P
from .capi import _ptr
from .hostmesh import NodalStorage


def taylor_green_ic(x, gamma=1.4, L=1.0, u0=1.0, rho0=1.0, p0=100.0):
    """UserDefinedInitialCondition of test/NavierStokes/TaylorGreen/SETUP/ProblemFile.f90:107-138.  x: [...,3]."""
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    rho = rho0 * np.ones_like(X)
    u = u0 * np.sin(X / L) * np.cos(Y / L) * np.cos(Z / L)
    v = -u0 * np.cos(X / L) * np.sin(Y / L) * np.cos(Z / L)
    w = np.zeros_like(X)
    p = p0 + rho0 / 16.0 * (np.cos(2.0 * X / L) * np.cos(2.0 * Z / L) + 2.0 * np.cos(2.0 * Y / L) + 2.0 * np.cos(2.0 * X / L)
                            + np.cos(2.0 * Y / L) * np.cos(2.0 * Z / L))
    Q = np.empty(x.shape[:-1] + (5,))
    Q[..., 0] = rho
    Q[..., 1] = rho * u
    Q[..., 2] = rho * v
    Q[..., 3] = rho * w
    Q[..., 4] = p / (gamma - 1.0) + 0.5 * rho * (u * u + v * v + w * w)
    return Q


class DGSem:
    def __init__(self, api, mesh, physics):
        self.api, self.mesh, self.physics = api, mesh, physics
        self.nElem, self.nFaces = mesh.nElem, mesh.nFaces
        self.mixed = bool(getattr(mesh, "mixed", False))
        api.set_physics(physics)
        if self.mixed:
            self._construct_mixed(api, mesh)
        else:
            self.N, self.n = mesh.N, mesh.N + 1
            self.NDOF = self.nElem * self.n ** 3            # nodes, as the reference counts them (main.f90:360)
            self.sp = NodalStorage(mesh.N, mesh.nodes)
            api.set_basis(self.sp)
            api.set_mesh(mesh)
            self._shape = (self.nElem, self.n, self.n, self.n, 5)
        bcs = getattr(mesh, "bcs", [])
        if bcs:
            types = [P.BC_TYPES[b[1].lower()] for b in bcs]
            params = mesh.bc_params if mesh.bc_params is not None else np.zeros((len(bcs), 16))
            api.set_boundary_conditions(types, params)
        if physics.les_wall_model:
            dw, dwf = np.ascontiguousarray(mesh.array("dWall")), np.ascontiguousarray(mesh.array("faceDWall"))
            if dw.size == 0:
                raise ValueError("the LES wall model needs wall distances: call HostMesh.wall_distances() first")
            if getattr(mesh, "is_partition", False) and not getattr(mesh, "wall_global", False):
                raise ValueError("wall distances of a partition must be measured against the wall nodes of every rank "
                                 "(HexMesh.f90:5594-5780): HostMesh.wall_distances(gather=...)")
            api.call("set_wall_distance", _ptr(dw, np.float64), _ptr(dwf, np.float64))
        if physics.viscous == P.VISCOUS["ip"]:
            api.call("set_face_h", _ptr(np.ascontiguousarray(mesh.array("faceH")), np.float64))
        counts = mesh.array("haloCount")
        if len(counts) and hasattr(api, "set_halo"):
            api.set_halo(mesh.array("haloRank"), counts, mesh.array("haloFace"), mesh.array("haloSide"))

    def _construct_mixed(self, api, mesh):
        """p-nonconforming mesh: NodalStorage(N) of every order in use (DGSEMClass.f90:215-228), Tset(N, M) of every
        pair of orders that meet at a face (Face_LinkWithElements, FaceClass.f90:236-251), then the mesh with its orders.  Element
        arrays are packed element after element at their own sizes, as the reference's global2LocalQ (StorageClass.f90:423-429)."""
        from .hostmesh import interpolation_matrix
        self.orders = np.array(mesh.array("elemOrder")).reshape(-1, 3)
        self.face_orders = np.array(mesh.array("faceOrder")).reshape(-1, 6)
        self.N = self.n = None
        self.sps = {int(N): NodalStorage(int(N), mesh.nodes) for N in sorted(set(self.orders.ravel()) | set(self.face_orders.ravel()))}
        for sp in self.sps.values():
            api.set_basis(sp)
        pairs = set()
        for s in (2, 4):
            for d in (0, 1):
                a, b = self.face_orders[:, s + d], self.face_orders[:, d]
                pairs |= {(int(x), int(y)) for x, y in zip(a[a != b], b[a != b])}
        for a, b in sorted(pairs):
            for No, Nd in ((a, b), (b, a)):
                T = interpolation_matrix(No, Nd, mesh.nodes)
                api.call("set_interpolation", No, Nd, _ptr(T, np.float64))
        api.set_mesh_p(mesh)
        sizes = np.prod(self.orders + 1, axis=1)
        self.elem_offset = np.concatenate([[0], np.cumsum(sizes)])
        self.NDOF = int(self.elem_offset[-1])
        self._shape = (self.NDOF, 5)

    def element_view(self, A, e):
        """Element e of a packed array of a p-nonconforming mesh, as [k][j][i][...]."""
        nx, ny, nz = self.orders[e] + 1
        return A[self.elem_offset[e]:self.elem_offset[e + 1]].reshape((nz, ny, nx) + A.shape[1:])

    # ---- state
    def node_coordinates(self):
        if self.mixed:
            return self.mesh.array("x").reshape(self.NDOF, 3)
        return self.mesh.array("x").reshape(self.nElem, self.n, self.n, self.n, 3)

    def set_Q(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(self._shape)
        self.api.call("upload_Q", _ptr(Q, np.float64))

    def set_initial_condition(self, fn, **kw):
        self.set_Q(fn(self.node_coordinates(), **kw))

    def set_source(self, S):
        if S is None:
            self.api.call("set_source", None)
        else:
            S = np.ascontiguousarray(S, dtype=np.float64).reshape(self._shape)
            self.api.call("set_source", _ptr(S, np.float64))

    def download(self, Q=False, QDot=False, gradients=False):
        out = {}
        bufs = [None] * 5
        names = ["Q", "QDot", "U_x", "U_y", "U_z"]
        want = [Q, QDot, gradients, gradients, gradients]
        for i, wnt in enumerate(want):
            if wnt:
                out[names[i]] = np.empty(self._shape)
                bufs[i] = _ptr(out[names[i]], np.float64)
        self.api.call("download", *bufs)
        return out

    def Q(self):
        return self.download(Q=True)["Q"]

    def QDot(self):
        return self.download(QDot=True)["QDot"]

    # ---- the reference's procedures
    def ComputeTimeDerivative(self, time=0.0):
        self.api.call("compute_time_derivative", float(time))

    # stage times t + b dt (ExplicitMethods.f90:690-692, 812-816, 903-905; d(k) of the SSP schemes :1002, :1128)
    _RK_B = {P.EULER: (0.0,), P.RK3: (0.0, 1.0 / 3.0, 3.0 / 4.0),
             P.RK5: (0.0, 0.1496590219993, 0.3704009573644, 0.6222557631345, 0.9582821306748),
             P.LSERK14_4: (0.0000000000000000, 0.0367762454319673, 0.1249685262725025, 0.2446177702277698, 0.2476149531070420, 0.2969311120382472,
                           0.3978149645802642, 0.5270854589440328, 0.6981269994175695, 0.8190890835352128, 0.8527059887098624, 0.8604711817462826,
                           0.8627060376969976, 0.8734213127600976),
             P.SSPRK33: (0.0, 1.0, 0.5), P.SSPRK43: (0.0, 0.5, 1.0, 0.5)}
    SCHEMES = {"euler": P.EULER, "rk3": P.RK3, "rk5": P.RK5, "lserk14-4": P.LSERK14_4, "ssprk33": P.SSPRK33, "ssprk43": P.SSPRK43}

    def _rk_step(self, scheme, t, dt, ctd_after_step, source):
        """source: None (the source set by set_source, constant over the step) or a callable time -> S array: the
        UserDefinedSourceTermNS of the reference, evaluated at every stage time (SpatialDiscretization.f90:569-577)."""
        if source is None:
            self.api.call("rk_step", scheme, float(t), float(dt), int(ctd_after_step))
            return
        for k, b in enumerate(self._RK_B[scheme]):
            self.set_source(source(t + b * dt))
            self.api.call("rk_stage", scheme, k, float(t), float(dt))
        if ctd_after_step:
            self.set_source(source(t + dt))
            self.ComputeTimeDerivative(t + dt)

    def TakeRK3Step(self, t, dt, ctd_after_step=False, source=None):
        self._rk_step(P.RK3, t, dt, ctd_after_step, source)

    def TakeRK5Step(self, t, dt, ctd_after_step=False, source=None):
        self._rk_step(P.RK5, t, dt, ctd_after_step, source)

    def TakeExplicitEulerStep(self, t, dt, ctd_after_step=False, source=None):
        self._rk_step(P.EULER, t, dt, ctd_after_step, source)

    def TakeLSERK14_4Step(self, t, dt, ctd_after_step=False, source=None):
        self._rk_step(P.LSERK14_4, t, dt, ctd_after_step, source)

    def TakeSSPRK33Step(self, t, dt, ctd_after_step=False, source=None):
        self._rk_step(P.SSPRK33, t, dt, ctd_after_step, source)

    def TakeSSPRK43Step(self, t, dt, ctd_after_step=False, source=None):
        self._rk_step(P.SSPRK43, t, dt, ctd_after_step, source)

    def MaxTimeStep(self, cfl, dcfl):
        a, b = C.c_double(), C.c_double()
        self.api.call("max_timestep", float(cfl), float(dcfl), C.byref(a), C.byref(b))
        return a.value, b.value

    def ComputeMaxResiduals(self):
        r = np.zeros(5)
        self.api.call("max_residuals", _ptr(r, np.float64))
        return r

    def ScalarVolumeIntegral(self, kind):
        v = C.c_double()
        self.api.call("volume_integral", int(kind), C.byref(v))
        return v.value

    def SurfaceIntegral(self, zone, kind):
        """ScalarSurfaceIntegral / VectorSurfaceIntegral (libs/monitors/SurfaceIntegrals.f90:40, 248); zone = name or index."""
        if isinstance(zone, str):
            zone = [b[0].lower() for b in self.mesh.bcs].index(zone.lower())
        out = np.zeros(3)
        self.api.call("surface_integral", int(zone), int(kind), _ptr(out, np.float64))
        return out[0] if kind <= P.SURF_PRESSURE else out

    def surface_monitor(self, zone, variable, direction=None, reference_surface=None, Lref=1.0):
        """SurfaceMonitor_Update (libs/monitors/SurfaceMonitor.f90:337-446) with the reference values of
        PhysicsStorage_NS.f90:290-306 (T_ref = 520 R, p_ref = 101325 Pa, R = 287.15)."""
        ph, variable = self.physics, variable.lower()
        T_ref = 520.0 * 5.0 / 9.0
        rho_ref = 101325.0 / (287.15 * T_ref)
        V_ref = ph.Mach * np.sqrt(ph.gamma * 287.15 * T_ref)
        d = None if direction is None else np.asarray(direction, dtype=np.float64)
        if variable == "mass-flow":
            return self.SurfaceIntegral(zone, P.SURF_MASS_FLOW)
        if variable == "flow":
            return self.SurfaceIntegral(zone, P.SURF_FLOW_RATE)
        if variable == "pressure-average":
            return self.SurfaceIntegral(zone, P.SURF_PRESSURE) / self.SurfaceIntegral(zone, P.SURF_SURFACE)
        if variable in ("pressure-force", "viscous-force", "force"):
            kind = {"pressure-force": P.SURF_PRESSURE_FORCE, "viscous-force": P.SURF_VISCOUS_FORCE, "force": P.SURF_TOTAL_FORCE}[variable]
            F = rho_ref * V_ref ** 2 * Lref ** 2 * self.SurfaceIntegral(zone, kind)
            return float(np.dot(F, d))
        if variable in ("lift", "drag"):
            kind = P.SURF_TOTAL_FORCE if ph.flowIsNavierStokes else P.SURF_PRESSURE_FORCE
            F = 2.0 * Lref ** 2 * self.SurfaceIntegral(zone, kind) / reference_surface
            return float(np.dot(F, d))
        raise ValueError("surface monitor variable not recognized: " + variable)

    def snapshot_begin(self):
        """Start an asynchronous copy of the current Q to the host (autosave beside the time loop)."""
        self.api.call("snapshot_begin")

    def snapshot_end(self):
        out = np.empty(self._shape)
        self.api.call("snapshot_end", _ptr(out, np.float64))
        return out

    def UpdateStatistics(self, reset=False):
        """StatisticsMonitor_UpdateValues (libs/monitors/StatisticsMonitor.f90:279)."""
        self.api.call("statistics_update", int(reset))

    def Statistics(self):
        """(data[e][k][j][i][var], samples): u v w uu vv ww uv uw vw, Q(5) [, U_x, U_y, U_z]."""
        nv, ns = C.c_int(), C.c_int()
        self.api.call("statistics_download", None, C.byref(nv), C.byref(ns))
        data = np.empty(self._shape[:-1] + (nv.value,))
        self.api.call("statistics_download", _ptr(data, np.float64), C.byref(nv), C.byref(ns))
        return data, ns.value

    def checkForNan(self):
        f = C.c_int()
        self.api.call("has_nan", C.byref(f))
        return bool(f.value)

    def volume_monitors(self):
        """VolumeMonitor_Update (libs/monitors/VolumeMonitor.f90:297-309)."""
        vol = self.ScalarVolumeIntegral(P.INT_VOLUME)
        return {
            "kinetic energy": self.ScalarVolumeIntegral(P.INT_KINETIC_ENERGY) / vol,
            "kinetic energy rate": self.ScalarVolumeIntegral(P.INT_KINETIC_ENERGY_RATE) / vol,
            "enstrophy": 0.5 * self.ScalarVolumeIntegral(P.INT_ENSTROPHY) / vol,
        }

    def enable_limiter(self, enabled=True, minimum=0.0):
        """Enable_limiter (ExplicitMethods.f90:1737-1752): positivity limiter after every SSPRK33 / SSPRK43 stage."""
        self.api.call("enable_limiter", int(bool(enabled)), float(minimum))

    VOLUME_MONITORS = {"kinetic energy": (P.INT_KINETIC_ENERGY, 1.0), "kinetic energy rate": (P.INT_KINETIC_ENERGY_RATE, 1.0),
                       "enstrophy": (P.INT_ENSTROPHY, 0.5), "entropy": (P.INT_ENTROPY, 1.0), "entropy rate": (P.INT_ENTROPY_RATE, 1.0),
                       "entropy balance": (P.INT_ENTROPY_BALANCE, 1.0), "math entropy": (P.INT_MATH_ENTROPY, 1.0),
                       "internal energy": (P.INT_INTERNAL_ENERGY, 1.0), "mean velocity": (P.INT_VELOCITY, 1.0),
                       "kinetic energy balance": (P.INT_KINETIC_ENERGY_BALANCE, 1.0)}

    def volume_monitor(self, variable):
        """One volume monitor by its control-file name (VolumeMonitor_Update, VolumeMonitor.f90:297-330)."""
        kind, factor = self.VOLUME_MONITORS[variable.lower()]
        return factor * self.ScalarVolumeIntegral(kind) / self.ScalarVolumeIntegral(P.INT_VOLUME)

    def integrate(self, nsteps, cfl=None, dcfl=None, dt=None, t0=0.0, scheme="RK3", monitors=True, t_final=None, source=None,
                  ctd_after_step=False, keep="all"):
        """Explicit branch of TimeIntegrator_t%integrate: initial residual, then per step
        MaxTimeStep -> CorrectDt -> RKStep -> ComputeMaxResiduals -> monitors (TimeIntegrator.f90:667-673, 737-959).
        t_final selects the time-accurate mode: the step is clipped to land on t_final (CorrectDt, :1130-1135) and the
        loop ends there (:882-887).  Returns the per-step records (the reference's monitor buffer lines), or only the
        last one with keep="last"."""
        code = self.SCHEMES[scheme.lower()]
        step = lambda t, dt, **kw: self._rk_step(code, t, dt, kw.get("ctd_after_step", False), kw.get("source"))
        t = t0
        if source is not None:
            self.set_source(source(t))
        self.ComputeTimeDerivative(t)
        rec = [dict(iter=0, t=t, dt=0.0, residuals=self.ComputeMaxResiduals(), **(self.volume_monitors() if monitors else {}))]
        eps = np.finfo(np.float64).eps
        for k in range(nsteps):
            if dt is None:
                dtc, dtv = self.MaxTimeStep(cfl, dcfl if dcfl is not None else cfl)
                step_dt = dtc if dtc < dtv else dtv       # DGSEMClass.f90:1025-1031
            else:
                step_dt = dt
            if t_final is not None and t + step_dt > t_final:
                step_dt = t_final - t
            step(t, step_dt, ctd_after_step=ctd_after_step, source=source)
            t = t + step_dt
            last = (t_final is not None and (t >= t_final or abs(t - t_final) <= 100.0 * eps)) or k == nsteps - 1
            if keep == "all" or last:
                if self.checkForNan():
                    raise FloatingPointError("Numerical divergence obtained in solver.")
                r = dict(iter=k + 1, t=t, dt=step_dt, residuals=self.ComputeMaxResiduals(), **(self.volume_monitors() if monitors else {}))
                rec = rec + [r] if keep == "all" else [r]
            if last:
                break
        return rec
