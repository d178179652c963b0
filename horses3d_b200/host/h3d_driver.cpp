// Native host-side driver above the C ABI: what NavierStokesSolver/main.f90 does around the hot path, for the cases the
// control-file keys below can express (the reference's control-file parser itself is out of scope).
//
//   h3d_driver --lib PATH [--prefix h3d_] [--mesh FILE | --ne 32 [--amp 0.1]] --order 3 | --order-file FILE [--nodes gauss|gauss-lobatto]
//              [--flow NS|Euler] [--mach 0.08] [--reynolds 1600] [--riemann roe] [--inviscid standard|split-form] [--averaging standard]
//              [--viscous BR1|BR2|IP] [--gradient-variables state|entropy|energy] [--les none|smagorinsky] [--lambda-stab 1]
//              [--bc name:type[:coupled]]... [--ic tgv|uniform] [--aoa-theta 0 --aoa-phi 0]
//              [--scheme rk3|rk5|euler|lserk14-4|ssprk33|ssprk43] [--steps 5] [--cfl 0.4 --dcfl 0.4 | --dt 1e-4] [--limiter MIN] [--device 0]
//
// Prints one monitor line per step (iteration, time, residuals, kinetic energy, its rate, enstrophy) as the reference's
// monitors do, and a final "FINAL ..." line in full precision.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>

#include "dgsem.hpp"

using namespace h3d;

static const double PI = 3.141592653589793238462643383279502884;

int main(int argc, char** argv) {
    std::map<std::string, std::string> opt = {{"prefix", "h3d_"}, {"ne", "8"}, {"amp", "0"}, {"order", "3"}, {"nodes", "gauss"}, {"flow", "NS"}, {"mach", "0.08"},
        {"reynolds", "1600"}, {"riemann", "roe"}, {"inviscid", "standard"}, {"averaging", "standard"}, {"viscous", "BR1"}, {"gradient-variables", "state"},
        {"les", "none"}, {"lambda-stab", "1"}, {"ic", "tgv"}, {"aoa-theta", "0"}, {"aoa-phi", "0"}, {"scheme", "rk3"}, {"steps", "5"}, {"cfl", "0.4"},
        {"dcfl", "0.4"}, {"dt", "0"}, {"device", "0"}, {"limiter", "0"}, {"lib", ""}, {"mesh", ""}, {"order-file", ""}};
    std::vector<std::string> bcArgs;
    for (int a = 1; a < argc; ++a) {
        std::string key = argv[a];
        if (key.rfind("--", 0) != 0 || a + 1 >= argc) { std::cerr << "usage: see the head of h3d_driver.cpp (bad argument " << key << ")\n"; return 2; }
        key = key.substr(2);
        if (key == "bc") { bcArgs.push_back(argv[++a]); continue; }
        if (!opt.count(key)) { std::cerr << "unknown option --" << key << "\n"; return 2; }
        opt[key] = argv[++a];
    }
    try {
        if (opt["lib"].empty()) throw std::runtime_error("--lib PATH (the shared library implementing include/h3d_gpu.h) is required");
        Backend api; api.load(opt["lib"], opt["prefix"]);
        DGSem sem(api);
        // ---- mesh (ConstructMeshFromFile / the generated periodic box) and boundary table (#define boundary ...)
        std::vector<BCSpec> bcs;
        std::string err;
        if (!opt["mesh"].empty()) {
            if (!readSpecMesh(opt["mesh"], sem.mesh, err)) throw std::runtime_error(err);
        } else {
            const int ne = std::atoi(opt["ne"].c_str());
            boxMesh(sem.mesh, ne, 2.0 * PI, std::atof(opt["amp"].c_str()), 2, 0, 1234u);
            if (bcArgs.empty()) bcArgs = {"front:periodic:back", "back:periodic:front", "bottom:periodic:top", "top:periodic:bottom", "left:periodic:right", "right:periodic:left"};
        }
        PhysicsOptions po;
        po.flow = opt["flow"]; po.mach = std::atof(opt["mach"].c_str()); po.reynolds = std::atof(opt["reynolds"].c_str()); po.riemann = opt["riemann"];
        po.inviscid = opt["inviscid"]; po.averaging = opt["averaging"]; po.viscous = opt["viscous"]; po.gradientVariables = opt["gradient-variables"];
        po.les = opt["les"]; po.lambdaStab = std::atof(opt["lambda-stab"].c_str());
        const H3dPhysics phys = makePhysics(po);
        const double theta = std::atof(opt["aoa-theta"].c_str()) * (PI / 180.0), phi = std::atof(opt["aoa-phi"].c_str()) * (PI / 180.0);
        const double uInf = std::cos(theta) * std::cos(phi), vInf = std::sin(theta) * std::cos(phi), wInf = std::sin(phi);
        for (auto& s : bcArgs) {
            BCSpec bc; size_t p1 = s.find(':'), p2 = s.find(':', p1 + 1);
            if (p1 == std::string::npos) throw std::runtime_error("--bc expects name:type[:coupled], got " + s);
            bc.name = toLower(s.substr(0, p1)); bc.type = toLower(s.substr(p1 + 1, p2 == std::string::npos ? std::string::npos : p2 - p1 - 1));
            bc.coupled = p2 == std::string::npos ? "" : toLower(s.substr(p2 + 1));
            std::memset(bc.params, 0, sizeof(bc.params));
            const double Tref = 520.0 * 5.0 / 9.0;
            if (bc.type == "noslipwall") bc.params[5] = Tref * phys.gammaM2 * phys.gammaMinus1;                 // NoSlipWallBC.f90:150-210 (adiabatic, at rest)
            else if (bc.type == "freeslipwall") bc.params[5] = Tref * phys.gammaM2;                             // FreeSlipWallBC.f90:150-200
            else if (bc.type == "inflow") {                                                                     // InflowBC.f90:150-330
                const double pIn = 1.0 / phys.gammaM2, rhoIn = 1.0, vIn = phys.Mach * std::sqrt(phys.gamma * pIn / rhoIn);
                bc.params[0] = rhoIn; bc.params[1] = vIn * uInf; bc.params[2] = vIn * vInf; bc.params[3] = vIn * wInf; bc.params[4] = pIn;
            } else if (bc.type == "outflow") bc.params[4] = 1.0 / phys.gammaM2;                                 // OutflowBC.f90:120-200
            bcs.push_back(bc);
        }
        if (!buildConnectivity(sem.mesh, bcs, err)) throw std::runtime_error(err);
        const int nodeType = toLower(opt["nodes"]) == "gauss" ? GAUSS : GAUSSLOBATTO;
        if (!opt["order-file"].empty()) {                              // "polynomial order file" (ReadOrderFile, ReadInputFile.f90:132-153)
            std::ifstream of(opt["order-file"]);
            if (!of) throw std::runtime_error("Error opening file: " + opt["order-file"]);
            int nelem = 0; of >> nelem;
            std::vector<int> orders(3 * (size_t)std::max(nelem, 0));
            for (auto& q : orders) if (!(of >> q)) throw std::runtime_error("bad polynomial order file");
            sem.constructP(orders, nodeType, phys, std::atoi(opt["device"].c_str()));
        } else
            sem.construct(std::atoi(opt["order"].c_str()), nodeType, phys, std::atoi(opt["device"].c_str()));
        // ---- UserDefinedInitialCondition
        if (toLower(opt["ic"]) == "tgv") {                             // test/NavierStokes/TaylorGreen/SETUP/ProblemFile.f90:107-138
            sem.setInitialCondition([&](const double* x, double* Q) {
                const double rho = 1.0, u = std::sin(x[0]) * std::cos(x[1]) * std::cos(x[2]), v = -std::cos(x[0]) * std::sin(x[1]) * std::cos(x[2]), w = 0.0;
                const double p = 100.0 + rho / 16.0 * (std::cos(2.0 * x[0]) * std::cos(2.0 * x[2]) + 2.0 * std::cos(2.0 * x[1]) + 2.0 * std::cos(2.0 * x[0])
                                                       + std::cos(2.0 * x[1]) * std::cos(2.0 * x[2]));
                Q[0] = rho; Q[1] = rho * u; Q[2] = rho * v; Q[3] = rho * w; Q[4] = p / (phys.gamma - 1.0) + 0.5 * rho * (u * u + v * v + w * w);
            });
        } else {                                                       // uniform flow at the angles of attack (test/NavierStokes/Cylinder, :304-322)
            sem.setInitialCondition([&](const double*, double* Q) {
                Q[0] = 1.0; Q[1] = uInf; Q[2] = vInf; Q[3] = wInf;
                Q[4] = (1.0 / phys.gammaM2) / (phys.gamma - 1.0) + 0.5 * (uInf * uInf + vInf * vInf + wInf * wInf);
            });
        }
        // ---- time integration
        TimeIntegrator ti;
        ti.scheme = lookup("explicit method", opt["scheme"], {{"rk3", H3D_RK3}, {"rk5", H3D_RK5}, {"euler", H3D_EULER}, {"lserk14-4", H3D_LSERK14_4},
                                                              {"ssprk33", H3D_SSPRK33}, {"ssprk43", H3D_SSPRK43}});
        ti.numberOfSteps = std::atoi(opt["steps"].c_str()); ti.cfl = std::atof(opt["cfl"].c_str()); ti.dcfl = std::atof(opt["dcfl"].c_str());
        ti.dt = std::atof(opt["dt"].c_str());
        if (std::atof(opt["limiter"].c_str()) > 0.0) sem.check(api.enable_limiter(sem.h, 1, std::atof(opt["limiter"].c_str())));
        std::printf("# elements %d faces %d order %d NDOF %lld\n", sem.mesh.nElem(), sem.mesh.nFaces, sem.N, sem.NDOF);
        ti.onStep = [](const MonitorLine& m) {
            std::printf("%6d %12.5e | %10.3e %10.3e %10.3e %10.3e %10.3e | %14.7e %14.7e %14.7e\n", m.iter, m.t, m.residuals[0], m.residuals[1],
                        m.residuals[2], m.residuals[3], m.residuals[4], m.kineticEnergy, m.kineticEnergyRate, m.enstrophy);
        };
        const MonitorLine last = ti.integrate(sem);
        std::printf("FINAL %d %.17e %.17e %.17e %.17e %.17e %.17e %.17e %.17e %.17e\n", last.iter, last.t, last.residuals[0], last.residuals[1], last.residuals[2],
                    last.residuals[3], last.residuals[4], last.kineticEnergy, last.kineticEnergyRate, last.enstrophy);
    } catch (const std::exception& ex) {
        std::cerr << "h3d_driver: " << ex.what() << "\n";
        return 1;
    }
    return 0;
}
