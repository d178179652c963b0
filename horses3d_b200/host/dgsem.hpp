// C++ mirror of the reference's driver objects above the C ABI of include/h3d_gpu.h:
//   Backend          the entry points of one shared library, bound at run time (what the Fortran adapter binds through ISO_C_BINDING)
//   makePhysics      ConstructPhysicsStorage_NS + SetRiemannSolver (PhysicsStorage_NS.f90:143-435, RiemannSolvers_NS.f90:120-286)
//   DGSem            DGSem (libs/discretization/DGSEMClass.f90:268-468): construct, ComputeMaxResiduals, MaxTimeStep
//   TimeIntegrator   TimeIntegrator_t % integrate, explicit branch (libs/timeintegrator/TimeIntegrator.f90:667-959)
// The library is named by path and symbol prefix, so the same driver runs on any implementation of the header.
#pragma once
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/h3d_gpu.h"
#include "geometry.hpp"
#include "geometry_p.hpp"

namespace h3d {

struct Backend {
    void* lib = nullptr; std::string prefix;
    template <class F> F sym(const char* name, bool required = true) {
        void* p = dlsym(lib, (prefix + name).c_str());
        if (!p && required) throw std::runtime_error("missing entry point " + prefix + name);
        return reinterpret_cast<F>(p);
    }
    int (*create)(h3d_handle*, int, int, int, const void*) = nullptr;
    int (*destroy)(h3d_handle) = nullptr;
    const char* (*last_error)(h3d_handle) = nullptr;
    int (*set_physics)(h3d_handle, const H3dPhysics*) = nullptr;
    int (*set_basis)(h3d_handle, int, int, const double*, const double*, const double*, const double*, const double*, const double*, const double*) = nullptr;
    int (*set_mesh)(h3d_handle, int, int, const int*, const int*, const int*, const int*, const int*, const int*, const int*, const double*, const double*,
                    const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*,
                    const double*) = nullptr;
    int (*set_interpolation)(h3d_handle, int, int, const double*) = nullptr;
    int (*set_mesh_p)(h3d_handle, int, int, const int*, const int*, const int*, const int*, const int*, const int*, const int*, const int*, const int*, const double*,
                      const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*,
                      const double*, const double*) = nullptr;
    int (*set_wall_distance)(h3d_handle, const double*, const double*) = nullptr;
    int (*set_face_h)(h3d_handle, const double*) = nullptr;
    int (*set_boundary_conditions)(h3d_handle, int, const int*, const double*) = nullptr;
    int (*upload_Q)(h3d_handle, const double*) = nullptr;
    int (*download)(h3d_handle, double*, double*, double*, double*, double*) = nullptr;
    int (*compute_time_derivative)(h3d_handle, double) = nullptr;
    int (*rk_step)(h3d_handle, int, double, double, int) = nullptr;
    int (*enable_limiter)(h3d_handle, int, double) = nullptr;
    int (*max_residuals)(h3d_handle, double*) = nullptr;
    int (*max_timestep)(h3d_handle, double, double, double*, double*) = nullptr;
    int (*volume_integral)(h3d_handle, int, double*) = nullptr;
    int (*surface_integral)(h3d_handle, int, int, double*) = nullptr;
    int (*has_nan)(h3d_handle, int*) = nullptr;

    void load(const std::string& path, const std::string& pfx) {
        lib = dlopen(path.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (!lib) throw std::runtime_error(std::string("cannot load ") + path + ": " + dlerror());
        prefix = pfx;
        create = sym<decltype(create)>("create_handle", false);          // an implementation whose plain create has another signature
        if (!create) create = sym<decltype(create)>("create");
        destroy = sym<decltype(destroy)>("destroy"); last_error = sym<decltype(last_error)>("last_error");
        set_physics = sym<decltype(set_physics)>("set_physics"); set_basis = sym<decltype(set_basis)>("set_basis"); set_mesh = sym<decltype(set_mesh)>("set_mesh");
        set_interpolation = sym<decltype(set_interpolation)>("set_interpolation"); set_mesh_p = sym<decltype(set_mesh_p)>("set_mesh_p");
        set_wall_distance = sym<decltype(set_wall_distance)>("set_wall_distance"); set_face_h = sym<decltype(set_face_h)>("set_face_h");
        set_boundary_conditions = sym<decltype(set_boundary_conditions)>("set_boundary_conditions");
        upload_Q = sym<decltype(upload_Q)>("upload_Q"); download = sym<decltype(download)>("download");
        compute_time_derivative = sym<decltype(compute_time_derivative)>("compute_time_derivative"); rk_step = sym<decltype(rk_step)>("rk_step");
        enable_limiter = sym<decltype(enable_limiter)>("enable_limiter");
        max_residuals = sym<decltype(max_residuals)>("max_residuals"); max_timestep = sym<decltype(max_timestep)>("max_timestep");
        volume_integral = sym<decltype(volume_integral)>("volume_integral"); surface_integral = sym<decltype(surface_integral)>("surface_integral");
        has_nan = sym<decltype(has_nan)>("has_nan");
    }
};

struct PhysicsOptions {
    std::string flow = "NS", inviscid = "standard", riemann = "roe", averaging = "standard", viscous = "BR1", gradientVariables = "state", les = "none";
    double mach = 0.08, reynolds = 1600.0, prandtl = 0.72, lambdaStab = 1.0, penalty = -1.0, smagorinskyCs = -1.0;
    int ipVariant = -1, lesWallModel = 0;
};

inline int lookup(const std::string& what, const std::string& key, const std::vector<std::pair<const char*, int>>& table) {
    const std::string k = toLower(key);
    for (auto& e : table) if (k == e.first) return e.second;
    throw std::runtime_error(what + " not recognized: " + key);
}

inline H3dPhysics makePhysics(const PhysicsOptions& o) {
    H3dPhysics p{};
    const double gamma = 1.4, gm1 = 1.4 - 1.0;                        // PhysicsStorage_NS.f90:118-120
    p.gamma = gamma; p.gammaMinus1 = gm1; p.Mach = o.mach; p.Pr = o.prandtl; p.Prt = o.prandtl;
    bool ns = toLower(o.flow) != "euler";
    if (ns) {
        p.Re = o.reynolds;
        if (o.reynolds != 0.0) { p.mu = 1.0 / o.reynolds; p.kappa = 1.0 / (gm1 * (o.mach * o.mach) * o.reynolds * o.prandtl); }   // :240-243
        p.mu_to_kappa = 1.0 / (gm1 * (o.mach * o.mach) * o.prandtl);                                                           // :250
    }
    p.les = lookup("LES model", o.les, {{"none", H3D_LES_NONE}, {"smagorinsky", H3D_LES_SMAGORINSKY}, {"wale", H3D_LES_WALE}, {"vreman", H3D_LES_VREMAN}});
    if (p.les != H3D_LES_NONE) ns = true;                             // :405-414
    p.flowIsNavierStokes = ns ? 1 : 0; p.computeGradients = ns ? 1 : 0;
    p.gammaM2 = gamma * (o.mach * o.mach);                            // :285
    const double Tref = 520.0 * 5.0 / 9.0, S = 198.6 * 5.0 / 9.0;      // :294, :419
    p.S_div_Tref = S / Tref; p.T_renorm = Tref / Tref;
    p.inviscid = lookup("inviscid discretization", o.inviscid, {{"standard", H3D_STANDARD_DG}, {"split-form", H3D_SPLIT_DG}});
    p.riemann = lookup("Riemann solver", o.riemann, {{"roe", H3D_RIEMANN_ROE}, {"lax-friedrichs", H3D_RIEMANN_LXF}, {"central", H3D_RIEMANN_CENTRAL},
                       {"rusanov", H3D_RIEMANN_RUSANOV}, {"standard roe", H3D_RIEMANN_STDROE}, {"u-diss", H3D_RIEMANN_UDISS}, {"roe-pike", H3D_RIEMANN_ROEPIKE},
                       {"low dissipation roe", H3D_RIEMANN_LOWDISSROE}, {"matrix dissipation", H3D_RIEMANN_MATRIXDISS}});
    p.averaging = lookup("averaging", o.averaging, {{"standard", H3D_AVG_STANDARD}, {"kennedy-gruber", H3D_AVG_KENNEDYGRUBER}, {"pirozzoli", H3D_AVG_PIROZZOLI},
                         {"ducros", H3D_AVG_DUCROS}, {"morinishi", H3D_AVG_MORINISHI}, {"entropy conserving", H3D_AVG_ENTROPYCONS}, {"chandrasekar", H3D_AVG_CHANDRASEKAR}});
    p.lambdaStab = p.riemann == H3D_RIEMANN_CENTRAL ? 0.0 : o.lambdaStab;       // RiemannSolvers_NS.f90:211-224
    p.viscous = lookup("viscous discretization", o.viscous, {{"br1", H3D_VISCOUS_BR1}, {"br2", H3D_VISCOUS_BR2}, {"ip", H3D_VISCOUS_IP}});
    p.penaltyParameter = o.penalty >= 0.0 ? o.penalty : (p.viscous == H3D_VISCOUS_BR2 ? 2.0 : (p.viscous == H3D_VISCOUS_IP ? 1.0 : 0.0));
    p.ipVariant = o.ipVariant;
    p.gradientVariables = ns ? lookup("gradient variables", o.gradientVariables, {{"state", H3D_GRADVARS_STATE}, {"entropy", H3D_GRADVARS_ENTROPY},
                                      {"energy", H3D_GRADVARS_ENERGY}}) : H3D_GRADVARS_STATE;
    // "LES model intensity" defaults: Smagorinsky 0.2, WALE 0.325, Vreman 0.07 (LESModels.f90:233-254, 337-356, 466-485)
    p.smagorinsky_Cs = o.smagorinskyCs >= 0.0 ? o.smagorinskyCs : (p.les == H3D_LES_WALE ? 0.325 : (p.les == H3D_LES_VREMAN ? 0.07 : 0.2));
    p.les_wall_model = o.lesWallModel;
    return p;
}

class DGSem {
public:
    Backend& api; h3d_handle h = nullptr;
    HostMesh mesh; HostGeometry geom; H3dPhysics phys{};
    HostGeometryP geomP; bool mixed = false;      // p-nonconforming mesh: element-wise polynomial orders (constructP)
    int N = 0, n = 0; long long NDOF = 0;

    explicit DGSem(Backend& b) : api(b) {}
    ~DGSem() { if (h) api.destroy(h); }
    void check(int rc) const { if (rc != 0) throw std::runtime_error(std::string("h3d: ") + (api.last_error(h) ? api.last_error(h) : "error")); }

    // sem % construct (DGSEMClass.f90:268-468): nodal storage, geometry, device set-up
    void construct(int order, int nodeType, const H3dPhysics& physics, int device = 0) {
        phys = physics; N = order; n = order + 1;
        buildGeometry(mesh, N, nodeType, geom);
        if (phys.les_wall_model) computeWallDistances(mesh, geom);
        NDOF = (long long)mesh.nElem() * n * n * n;                   // nodes, as main.f90:360
        int rc = api.create(&h, 0, 1, device, nullptr);
        if (rc != 0) throw std::runtime_error(std::string("h3d_create failed: ") + (api.last_error(nullptr) ? api.last_error(nullptr) : "?"));
        check(api.set_physics(h, &phys));
        const NodalStorage& sp = geom.sp;
        check(api.set_basis(h, N, nodeType, sp.x.data(), sp.w.data(), sp.D.data(), sp.hatD.data(), sp.sharpD.data(), sp.v.data(), sp.b.data()));
        check(api.set_mesh(h, mesh.nElem(), mesh.nFaces, mesh.elemFace.data(), mesh.elemFaceSide.data(), mesh.faceElem.data(), mesh.faceElemSide.data(),
                           mesh.faceRot.data(), mesh.faceType.data(), mesh.faceZone.data(), geom.jGradXi.data(), geom.jGradEta.data(), geom.jGradZeta.data(),
                           geom.jac.data(), geom.x.data(), geom.volume.data(), geom.fnormal.data(), geom.ft1.data(), geom.ft2.data(), geom.fjac.data(),
                           geom.fx.data(), geom.fsurface.data()));
        setBoundaryTable();
        if (phys.les_wall_model) check(api.set_wall_distance(h, geom.dWall.data(), geom.fdWall.data()));
        if (phys.viscous == H3D_VISCOUS_IP) check(api.set_face_h(h, geom.fh.data()));
    }

    void setBoundaryTable() {
        if (mesh.bcs.empty()) return;
        std::vector<int> types; std::vector<double> params;
        for (auto& bc : mesh.bcs) {
            types.push_back(lookup("boundary condition", bc.type, {{"periodic", H3D_BC_PERIODIC}, {"noslipwall", H3D_BC_NOSLIPWALL},
                                   {"freeslipwall", H3D_BC_FREESLIPWALL}, {"inflow", H3D_BC_INFLOW}, {"outflow", H3D_BC_OUTFLOW}}));
            params.insert(params.end(), bc.params, bc.params + 16);
        }
        check(api.set_boundary_conditions(h, (int)types.size(), types.data(), params.data()));
    }

    // sem % construct with a polynomial order file (DGSEMClass.f90:195-228, pAdaptationClass.f90:218-220): orders[e][3]; nodal storages of
    // every order, interpolation matrices of every pair of orders that meet at a face (FaceClass.f90:236-251), h3d_set_mesh_p
    void constructP(const std::vector<int>& orders, int nodeType, const H3dPhysics& physics, int device = 0) {
        phys = physics; mixed = true; N = -1; n = 0;
        if ((int)orders.size() != 3 * mesh.nElem()) throw std::runtime_error("the polynomial order file does not match the number of elements of the mesh");
        std::string err;
        if (!buildGeometryP(mesh, orders.data(), nodeType, geomP, err)) throw std::runtime_error(err);
        NDOF = geomP.eOff[mesh.nElem()];
        int rc = api.create(&h, 0, 1, device, nullptr);
        if (rc != 0) throw std::runtime_error(std::string("h3d_create failed: ") + (api.last_error(nullptr) ? api.last_error(nullptr) : "?"));
        check(api.set_physics(h, &phys));
        for (auto& kv : geomP.sp) {
            const NodalStorage& sp = kv.second;
            check(api.set_basis(h, sp.N, nodeType, sp.x.data(), sp.w.data(), sp.D.data(), sp.hatD.data(), sp.sharpD.data(), sp.v.data(), sp.b.data()));
        }
        std::vector<double> T;
        for (int f = 0; f < mesh.nFaces; ++f) for (int s = 2; s <= 4; s += 2) for (int d = 0; d < 2; ++d) {
            const int a = geomP.faceOrder[6 * f + s + d], b = geomP.faceOrder[6 * f + d];
            if (a == b) continue;
            interpolationMatrix(geomP.sp.at(a), geomP.sp.at(b), T); check(api.set_interpolation(h, a, b, T.data()));
            interpolationMatrix(geomP.sp.at(b), geomP.sp.at(a), T); check(api.set_interpolation(h, b, a, T.data()));
        }
        check(api.set_mesh_p(h, mesh.nElem(), mesh.nFaces, geomP.elemOrder.data(), geomP.faceOrder.data(), mesh.elemFace.data(), mesh.elemFaceSide.data(), mesh.faceElem.data(),
                             mesh.faceElemSide.data(), mesh.faceRot.data(), mesh.faceType.data(), mesh.faceZone.data(), geomP.jGradXi.data(), geomP.jGradEta.data(),
                             geomP.jGradZeta.data(), geomP.jac.data(), geomP.x.data(), geomP.volume.data(), geomP.fnormal.data(), geomP.ft1.data(), geomP.ft2.data(),
                             geomP.fjac.data(), geomP.fx.data(), geomP.fsurface.data()));
        setBoundaryTable();
        if (phys.viscous == H3D_VISCOUS_IP) check(api.set_face_h(h, geomP.fh.data()));
        if (phys.les_wall_model) {
            computeWallDistancesP(geomP, wallCoordinatesP(mesh, geomP));
            check(api.set_wall_distance(h, geomP.dWall.data(), geomP.fdWall.data()));
        }
    }

    // UserDefinedInitialCondition: fn(x[3], Q[5]) at every node
    void setInitialCondition(const std::function<void(const double*, double*)>& fn) {
        std::vector<double> Q((size_t)NDOF * 5);
        const std::vector<double>& xn = mixed ? geomP.x : geom.x;
#pragma omp parallel for schedule(static)
        for (long long g = 0; g < NDOF; ++g) fn(&xn[3 * g], &Q[5 * g]);
        check(api.upload_Q(h, Q.data()));
    }
    std::vector<double> Q() { std::vector<double> q((size_t)NDOF * 5); check(api.download(h, q.data(), nullptr, nullptr, nullptr, nullptr)); return q; }

    void ComputeTimeDerivative(double t) { check(api.compute_time_derivative(h, t)); }
    void ComputeMaxResiduals(double r[5]) { check(api.max_residuals(h, r)); }
    double MaxTimeStep(double cfl, double dcfl) {                     // DGSEMClass.f90:1025-1031
        double dtc, dtv; check(api.max_timestep(h, cfl, dcfl, &dtc, &dtv));
        return dtc < dtv ? dtc : dtv;
    }
    double ScalarVolumeIntegral(int kind) { double v; check(api.volume_integral(h, kind, &v)); return v; }
    bool checkForNan() { int f; check(api.has_nan(h, &f)); return f != 0; }
};

struct MonitorLine { int iter; double t, dt, residuals[5], kineticEnergy, kineticEnergyRate, enstrophy; };

// TimeIntegrator_t % integrate, explicit branch: initial residual, then per step MaxTimeStep -> CorrectDt -> RKStep ->
// ComputeMaxResiduals -> monitors -> checkForNan (TimeIntegrator.f90:667-673, 737-959, 1100-1135)
class TimeIntegrator {
public:
    int scheme = H3D_RK3, numberOfSteps = 0; double cfl = 0.0, dcfl = 0.0, dt = 0.0, tFinal = -1.0; bool ctdAfterStep = false, volumeMonitors = true;
    std::function<void(const MonitorLine&)> onStep;

    MonitorLine monitors(DGSem& sem, int iter, double t, double stepDt) const {
        MonitorLine m{}; m.iter = iter; m.t = t; m.dt = stepDt;
        sem.ComputeMaxResiduals(m.residuals);
        if (volumeMonitors) {                                          // VolumeMonitor_Update (VolumeMonitor.f90:297-309)
            const double vol = sem.ScalarVolumeIntegral(H3D_INT_VOLUME);
            m.kineticEnergy = sem.ScalarVolumeIntegral(H3D_INT_KINETIC_ENERGY) / vol;
            m.kineticEnergyRate = sem.ScalarVolumeIntegral(H3D_INT_KINETIC_ENERGY_RATE) / vol;
            m.enstrophy = sem.phys.computeGradients ? 0.5 * sem.ScalarVolumeIntegral(H3D_INT_ENSTROPHY) / vol : 0.0;
        }
        return m;
    }

    MonitorLine integrate(DGSem& sem, double t0 = 0.0) const {
        double t = t0;
        sem.ComputeTimeDerivative(t);
        MonitorLine last = monitors(sem, 0, t, 0.0);
        if (onStep) onStep(last);
        const double eps = std::numeric_limits<double>::epsilon();
        for (int k = 0; k < numberOfSteps; ++k) {
            double stepDt = dt > 0.0 ? dt : sem.MaxTimeStep(cfl, dcfl > 0.0 ? dcfl : cfl);
            if (tFinal >= 0.0 && t + stepDt > tFinal) stepDt = tFinal - t;        // CorrectDt
            sem.check(sem.api.rk_step(sem.h, scheme, t, stepDt, ctdAfterStep ? 1 : 0));
            t = t + stepDt;
            if (sem.checkForNan()) throw std::runtime_error("Numerical divergence obtained in solver.");   // ExplicitMethods.f90:1893-1905
            last = monitors(sem, k + 1, t, stepDt);
            if (onStep) onStep(last);
            if (tFinal >= 0.0 && (t >= tFinal || std::fabs(t - tFinal) <= 100.0 * eps)) break;
        }
        return last;
    }
};

}  // namespace h3d
