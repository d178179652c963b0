// Host side of h3d_set_physics: validation of the run-time physics (the reference's error messages) and the trimmed kernel
// parameter made from it.  Shared by libh3dgpu.so (h3d_api.cu) and the host-loop backend of tests/emu.
#pragma once
#include <string>

#include "h3d_physics.cuh"

namespace h3d {

inline int physFromH3dPhysics(const H3dPhysics* p, Phys& q, std::string& err) {
    if (p->riemann < H3D_RIEMANN_ROE || p->riemann > H3D_RIEMANN_MATRIXDISS) { err = "Riemann Solver not recognized."; return 1; }
    if (p->averaging < H3D_AVG_STANDARD || p->averaging > H3D_AVG_CHANDRASEKAR) { err = "Averaging not recognized."; return 1; }
    if (p->inviscid != H3D_STANDARD_DG && p->inviscid != H3D_SPLIT_DG) { err = "Requested inviscid discretization is not implemented."; return 1; }
    if (p->viscous < H3D_VISCOUS_BR1 || p->viscous > H3D_VISCOUS_IP) { err = "Requested viscous discretization is not implemented."; return 1; }
    if (p->ipVariant < -1 || p->ipVariant > 1) { err = "Unknown selected IP variant."; return 1; }
    if (p->gradientVariables < H3D_GRADVARS_STATE || p->gradientVariables > H3D_GRADVARS_ENERGY) { err = "Gradient variables are not currently implemented."; return 1; }
    if (p->les < H3D_LES_NONE || p->les > H3D_LES_VREMAN) { err = "LES model not recognized."; return 1; }
    if (p->les_wall_model != 0 && p->les_wall_model != 1) { err = "LES wall model not recognized."; return 1; }
    q.gamma = p->gamma; q.gm1 = p->gammaMinus1; q.gammaM2 = p->gammaM2; q.mu = p->mu; q.mu_to_kappa = p->mu_to_kappa;
    q.S_div_Tref = p->S_div_Tref; q.T_renorm = p->T_renorm; q.lambdaStab = p->lambdaStab; q.Cs = p->smagorinsky_Cs;
    q.ns = p->flowIsNavierStokes; q.riemann = p->riemann; q.averaging = p->averaging; q.les = p->les;
    q.viscous = p->flowIsNavierStokes ? p->viscous : H3D_VISCOUS_BR1; q.ipVariant = p->ipVariant; q.eta = p->penaltyParameter;
    q.gradVars = p->flowIsNavierStokes ? p->gradientVariables : H3D_GRADVARS_STATE;
    q.wallModel = (p->les != H3D_LES_NONE && p->les_wall_model == 1) ? 1 : 0;
    return 0;
}

}  // namespace h3d
