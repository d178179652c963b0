// Per-node device functions of the compressible Navier-Stokes path (FP64).
//
// Each function states the reference routine whose arithmetic it reproduces (paths relative to
// /root/reference/Solver/src/libs/physics/navierstokes unless noted).  The operation order inside an
// expression follows the reference so that a build with -fmad=false is bit-comparable with the oracle.
#pragma once
#include <cuda_runtime.h>

#include "h3d_gpu.h"

namespace h3d {

struct Phys {   // by-value kernel parameter (a trimmed H3dPhysics)
    double gamma, gm1, gammaM2, mu, mu_to_kappa, S_div_Tref, T_renorm, lambdaStab, Cs;
    double eta;          // BR2 eta (EllipticBR2.f90:80-91)
    double penaltyNum;   // IP: 0.5 * sigma * (N+1) * (N+2), divided by the face's h in the kernel (EllipticIP.f90:678-687)
    int ns, riemann, averaging, les, wallModel;
    int viscous, ipVariant;   // H3D_VISCOUS_*, IP variant -1 / 0 / 1
    int gradVars;             // H3D_GRADVARS_*
};

__device__ __forceinline__ double pow2(double x) { return x * x; }

// Physics_NS.f90:51-97 (EulerFlux); F[c][d]
__device__ __forceinline__ void euler_flux(const Phys& ph, const double Q[5], double F[5][3]) {
    const double u = Q[1] / Q[0], v = Q[2] / Q[0], w = Q[3] / Q[0];
    const double p = ph.gm1 * (Q[4] - 0.5 * (Q[1] * u + Q[2] * v + Q[3] * w));
    F[0][0] = Q[1]; F[1][0] = Q[1] * u + p; F[2][0] = Q[1] * v; F[3][0] = Q[1] * w; F[4][0] = (Q[4] + p) * u;
    F[0][1] = Q[2]; F[1][1] = F[2][0]; F[2][1] = Q[2] * v + p; F[3][1] = Q[2] * w; F[4][1] = (Q[4] + p) * v;
    F[0][2] = Q[3]; F[1][2] = F[3][0]; F[2][2] = F[3][1]; F[3][2] = Q[3] * w + p; F[4][2] = (Q[4] + p) * w;
}

// VariableConversion_NS.f90:50-64,99-115,147-186: mu and kappa from the state (Sutherland)
__device__ __forceinline__ double pressure(const Phys& ph, const double Q[5]) {
    return ph.gm1 * (Q[4] - 0.5 * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]) / Q[0]);
}
__device__ __forceinline__ double sutherland(const Phys& ph, double T) {
    const double tT = T * ph.T_renorm;
    return (1.0 + ph.S_div_Tref) / (tT + ph.S_div_Tref) * tT * sqrt(tT);
}
__device__ __forceinline__ void laminar_mu_kappa(const Phys& ph, const double Q[5], double& mu, double& kappa) {
    const double T = ph.gammaM2 * pressure(ph, Q) / Q[0];
    mu = ph.mu * sutherland(ph, T);
    kappa = mu * ph.mu_to_kappa;
}

// VariableConversion_NS.f90:373-392
__device__ __forceinline__ void velocity_gradients(const double Q[5], const double Qx[5], const double Qy[5], const double Qz[5],
                                                   double ux[3], double uy[3], double uz[3]) {
    const double invRho = 1.0 / Q[0], invRho2 = invRho * invRho;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double uDivRho = Q[1 + c] * invRho2;
        ux[c] = invRho * Qx[1 + c] - uDivRho * Qx[0];
        uy[c] = invRho * Qy[1 + c] - uDivRho * Qy[0];
        uz[c] = invRho * Qz[1 + c] - uDivRho * Qz[0];
    }
}

// GetGradients procedure pointer: NSGradientVariables_STATE / _ENTROPY / _ENERGY (VariableConversion_NS.f90:196-262)
__device__ __forceinline__ void get_gradients(const Phys& ph, const double Q[5], double U[5]) {
    if (ph.gradVars == H3D_GRADVARS_ENTROPY) {
        const double invRho = 1.0 / Q[0];
        const double rhoV2 = (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) * invRho;
        const double p = ph.gm1 * (Q[4] - 0.5 * rhoV2);
        const double invP = 1.0 / p;
        const double invGm1 = 1.0 / ph.gm1;
        const double U0 = (ph.gamma - (log(p) - ph.gamma * log(Q[0]))) * invGm1 - 0.5 * rhoV2 * invP;
        const double U4 = -Q[0] * invP;
        U[1] = Q[1] * invP; U[2] = Q[2] * invP; U[3] = Q[3] * invP; U[0] = U0; U[4] = U4;
    } else if (ph.gradVars == H3D_GRADVARS_ENERGY) {
        const double invRho = 1.0 / Q[0];
        const double rhoV2 = (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) * invRho;
        const double p = ph.gm1 * (Q[4] - 0.5 * rhoV2);
        const double U4 = ph.gammaM2 * p * invRho, U0 = Q[0];
        U[1] = Q[1] * invRho; U[2] = Q[2] * invRho; U[3] = Q[3] * invRho; U[0] = U0; U[4] = U4;
    } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) U[q] = Q[q];
    }
}

// getVelocityGradients procedure pointer (VariableConversion_NS.f90:373-428, set at :602-617); GV = false: State only
template <bool GV>
__device__ __forceinline__ void velocity_gradients_gv(const Phys& ph, const double Q[5], const double Qx[5], const double Qy[5], const double Qz[5],
                                                      double ux[3], double uy[3], double uz[3]) {
    if (GV && ph.gradVars == H3D_GRADVARS_ENERGY) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { ux[c] = Qx[1 + c]; uy[c] = Qy[1 + c]; uz[c] = Qz[1 + c]; }
    } else if (GV && ph.gradVars == H3D_GRADVARS_ENTROPY) {   // as written in the reference (:421-426)
        const double pDivRho = pressure(ph, Q) / Q[0];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double U = Q[1 + c] / Q[0];
            ux[c] = pDivRho * Qx[1 + c] + U / pDivRho * Qx[4];
            uy[c] = pDivRho * Qy[1 + c] + U / pDivRho * Qy[4];
            uz[c] = pDivRho * Qz[1 + c] + U / pDivRho * Qz[4];
        }
    } else {
        velocity_gradients(Q, Qx, Qy, Qz, ux, uy, uz);
    }
}

// velocity gradients of getStressTensor (Physics_NS.f90:840-866): its entropy branch differs from the pointer's
__device__ __forceinline__ void stress_velocity_gradients(const Phys& ph, const double Q[5], const double Qx[5], const double Qy[5], const double Qz[5],
                                                          double ux[3], double uy[3], double uz[3]) {
    if (ph.gradVars == H3D_GRADVARS_ENTROPY) {
        const double invRho = 1.0 / Q[0];
        const double pdr = ph.gm1 * invRho * (Q[4] - 0.5 * (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) * invRho);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double u = Q[1 + c] * invRho;
            ux[c] = pdr * (Qx[1 + c] + u * Qx[4]); uy[c] = pdr * (Qy[1 + c] + u * Qy[4]); uz[c] = pdr * (Qz[1 + c] + u * Qz[4]);
        }
    } else {
        velocity_gradients_gv<true>(ph, Q, Qx, Qy, Qz, ux, uy, uz);
    }
}

// libs/physics/common/LESModels.f90:256-305 Smagorinsky: mu_t = rho LS^2 sqrt(2 S:S), S summed column by column;
// LS = Cs delta, limited to 0.4 dWall by the linear wall model (LESModel_ComputeWallEffect, :189-203)
// WALE_ComputeViscosity (LESModels.f90:358-435) and Vreman_ComputeViscosity (:487-546); only in the general instantiations
__device__ __noinline__ double wale_vreman(const Phys& ph, double delta, const double Q[5], const double ux[3], const double uy[3], const double uz[3]) {
    const double gradV[3][3] = {{ux[0], ux[1], ux[2]}, {uy[0], uy[1], uy[2]}, {uz[0], uz[1], uz[2]}};
    if (ph.les == H3D_LES_WALE) {
        double S[3][3], g2[3][3], Sd[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                S[i][j] = 0.5 * (gradV[i][j] + gradV[j][i]);
                g2[i][j] = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) g2[i][j] = g2[i][j] + gradV[i][k] * gradV[k][j];
            }
        const double divV2 = g2[0][0] + g2[1][1] + g2[2][2];
        double normS = 0.0, normSd = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) normS = normS + S[i][j] * S[i][j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Sd[i][j] = 0.5 * (g2[i][j] + g2[j][i]);
        Sd[0][0] = Sd[0][0] - 1.0 / 3.0 * divV2; Sd[1][1] = Sd[1][1] - 1.0 / 3.0 * divV2; Sd[2][2] = Sd[2][2] - 1.0 / 3.0 * divV2;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) normSd = normSd + Sd[i][j] * Sd[i][j];
        const double LS = ph.Cs * delta;
        double mu = Q[0] * pow2(LS) * (pow(normSd, 3.0 / 2.0) / (pow(normS, 5.0 / 2.0) + pow(normSd, 5.0 / 4.0)));
        if (normS < 1.0e-8 && normSd < 1.0e-8) mu = 0.0;
        return mu;
    }
    const double delta2 = delta * delta;
    double G[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            G[i][j] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) G[i][j] = G[i][j] + (gradV[i][k] * gradV[j][k] * delta2);
        }
    double alpha = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) alpha = alpha + gradV[i][j] * gradV[i][j];
    const double Bbeta = G[0][0] * G[1][1] + G[1][1] * G[2][2] + G[2][2] * G[0][0] - G[0][1] * G[0][1] - G[1][2] * G[1][2] - G[0][2] * G[0][2];
    return alpha > 1.0e-10 ? Q[0] * ph.Cs * sqrt(fabs(Bbeta) / alpha) : 0.0;
}

// LESModel % ComputeViscosity: Smagorinsky everywhere, WALE and Vreman in the general (GV) instantiations
template <bool GV = false>
__device__ __forceinline__ double smagorinsky(const Phys& ph, double delta, double dWall, const double Q[5], const double Qx[5], const double Qy[5], const double Qz[5]) {
    double ux[3], uy[3], uz[3];
    velocity_gradients_gv<GV>(ph, Q, Qx, Qy, Qz, ux, uy, uz);
    if (GV && ph.les != H3D_LES_SMAGORINSKY) return wale_vreman(ph, delta, Q, ux, uy, uz);
    // S(i,j) = 1/2 (column j of grad u + row contribution), built exactly as the reference does
    const double s00 = 0.5 * (ux[0] + ux[0]), s10 = 0.5 * (ux[1] + uy[0]), s20 = 0.5 * (ux[2] + uz[0]);
    const double s01 = 0.5 * (uy[0] + ux[1]), s11 = 0.5 * (uy[1] + uy[1]), s21 = 0.5 * (uy[2] + uz[1]);
    const double s02 = 0.5 * (uz[0] + ux[2]), s12 = 0.5 * (uz[1] + uy[2]), s22 = 0.5 * (uz[2] + uz[2]);
    double sum = 0.0;
    sum = sum + s00 * s00; sum = sum + s10 * s10; sum = sum + s20 * s20;
    sum = sum + s01 * s01; sum = sum + s11 * s11; sum = sum + s21 * s21;
    sum = sum + s02 * s02; sum = sum + s12 * s12; sum = sum + s22 * s22;
    const double normS = sqrt(2.0 * sum);
    double LS = ph.Cs * delta;
    if (ph.wallModel) LS = fmin(LS, dWall * 0.4);
    return Q[0] * pow2(LS) * normS;
}

// Physics_NS.f90:246-304 (ViscousFlux_STATE), :306-359 (_ENTROPY), :361-414 (_ENERGY); beta = 0 kept as in the reference
// call sites.  GV = false: State gradient variables only (the headline instantiations); GV = true: run-time choice.
template <bool GV = false>
__device__ __forceinline__ void viscous_flux(const Phys& ph, const double Q[5], const double Qx[5], const double Qy[5], const double Qz[5],
                                             double mu, double beta, double kappa, double F[5][3]) {
    const double invRho = 1.0 / Q[0];
    const double u = Q[1] * invRho, v = Q[2] * invRho, w = Q[3] * invRho;
    double ux[3], uy[3], uz[3], Tx, Ty, Tz;
    if (GV && ph.gradVars == H3D_GRADVARS_ENTROPY) {
        const double pdr = ph.gm1 * invRho * (Q[4] - 0.5 * (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) * invRho);
        const double uu[3] = {u, v, w};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ux[c] = pdr * (Qx[1 + c] + uu[c] * Qx[4]); uy[c] = pdr * (Qy[1 + c] + uu[c] * Qy[4]); uz[c] = pdr * (Qz[1 + c] + uu[c] * Qz[4]);
        }
        Tx = ph.gammaM2 * pow2(pdr) * Qx[4]; Ty = ph.gammaM2 * pow2(pdr) * Qy[4]; Tz = ph.gammaM2 * pow2(pdr) * Qz[4];
    } else if (GV && ph.gradVars == H3D_GRADVARS_ENERGY) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { ux[c] = Qx[1 + c]; uy[c] = Qy[1 + c]; uz[c] = Qz[1 + c]; }
        Tx = Qx[4]; Ty = Qy[4]; Tz = Qz[4];
    } else {
        const double uDivRho[3] = {u * invRho, v * invRho, w * invRho};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ux[c] = invRho * Qx[1 + c] - uDivRho[c] * Qx[0];
            uy[c] = invRho * Qy[1 + c] - uDivRho[c] * Qy[0];
            uz[c] = invRho * Qz[1 + c] - uDivRho[c] * Qz[0];
        }
        const double c0 = ph.gm1 * ph.gammaM2;
        Tx = c0 * (invRho * Qx[4] - Q[4] * invRho * invRho * Qx[0] - u * ux[0] - v * ux[1] - w * ux[2]);
        Ty = c0 * (invRho * Qy[4] - Q[4] * invRho * invRho * Qy[0] - u * uy[0] - v * uy[1] - w * uy[2]);
        Tz = c0 * (invRho * Qz[4] - Q[4] * invRho * invRho * Qz[0] - u * uz[0] - v * uz[1] - w * uz[2]);
    }
    const double divV = ux[0] + uy[1] + uz[2];
    F[0][0] = 0.0;
    F[1][0] = mu * (2.0 * ux[0] - 2.0 / 3.0 * divV) + beta * divV;
    F[2][0] = mu * (ux[1] + uy[0]);
    F[3][0] = mu * (ux[2] + uz[0]);
    F[4][0] = F[1][0] * u + F[2][0] * v + F[3][0] * w + kappa * Tx;
    F[0][1] = 0.0;
    F[1][1] = F[2][0];
    F[2][1] = mu * (2.0 * uy[1] - 2.0 / 3.0 * divV) + beta * divV;
    F[3][1] = mu * (uy[2] + uz[1]);
    F[4][1] = F[1][1] * u + F[2][1] * v + F[3][1] * w + kappa * Ty;
    F[0][2] = 0.0;
    F[1][2] = F[3][0];
    F[2][2] = F[3][1];
    F[3][2] = mu * (2.0 * uz[2] - 2.0 / 3.0 * divV) + beta * divV;
    F[4][2] = F[1][2] * u + F[2][2] * v + F[3][2] * w + kappa * Tz;
}

// libs/foundation/Utilities.f90:269-300 (logarithmicMean, Ismail & Roe)
__device__ __forceinline__ double log_mean(double aL, double aR) {
    const double xi = aL / aR;
    const double f = (xi - 1.0) / (xi + 1.0);
    const double u = f * f;
    double FF;
    if (u < 0.01) FF = 1.0 + (1.0 / 3.0) * u + (1.0 / 5.0) * pow2(u) + (1.0 / 7.0) * (u * u * u);
    else FF = log(xi) / (2.0 * f);
    return 0.5 * (aL + aR) / FF;
}

// Ismail-Roe (entropy conserving) and Chandrasekar mean states shared by the averages and the two-point fluxes
// (RiemannSolvers_NS.f90:1964-2077 and :2439-2641; the two families differ in 1/2 (zL+zR) versus (zL+zR) only)
__device__ __forceinline__ void ec_mean_state(const Phys& ph, bool twoPoint, double rhoL, double rhoR, double uL, double uR, double vL, double vR,
                                              double wL, double wR, double pL, double pR, double& rho, double& u, double& v, double& w, double& p, double& h) {
    const double gamma = ph.gamma, gm1 = ph.gm1;
    const double gammaPlus1Div2 = (gamma + 1.0) / 2.0, gammaMinus1Div2 = gm1 / 2.0, gammaDivGammaMinus1 = gamma / gm1, invGamma = 1.0 / gamma;
    double zL[5], zR[5], z[5];
    zL[4] = sqrt(rhoL * pL); zR[4] = sqrt(rhoR * pR);
    zL[0] = rhoL / zL[4]; zR[0] = rhoR / zR[4];
    zL[1] = zL[0] * uL; zR[1] = zR[0] * uR;
    zL[2] = zL[0] * vL; zR[2] = zR[0] * vR;
    zL[3] = zL[0] * wL; zR[3] = zR[0] * wR;
    const double z1Log = log_mean(zL[0], zR[0]), z5Log = log_mean(zL[4], zR[4]);
    if (twoPoint) {
#pragma unroll
        for (int q = 0; q < 5; ++q) z[q] = 0.5 * (zL[q] + zR[q]);
        rho = z[0] * z5Log;
    } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) z[q] = zL[q] + zR[q];
        rho = 0.5 * z[0] * z5Log;
    }
    const double invZ1 = 1.0 / z[0];
    u = z[1] * invZ1; v = z[2] * invZ1; w = z[3] * invZ1; p = z[4] * invZ1;
    const double p2 = (gammaPlus1Div2 * z5Log / z1Log + gammaMinus1Div2 * p) * invGamma;
    h = gammaDivGammaMinus1 * p2 / rho + 0.5 * (pow2(u) + pow2(v) + pow2(w));
}
// betaL, betaR = 1/2 rho / p of the two states (evaluated per pair by chandrasekar_mean_state, once per node by the volume kernel)
__device__ __forceinline__ void chandrasekar_mean_state_beta(const Phys& ph, double rhoL, double rhoR, double uL, double uR, double vL, double vR,
                                                             double wL, double wR, double betaL, double betaR, double& rho, double& u, double& v, double& w, double& p, double& h);
__device__ __forceinline__ void chandrasekar_mean_state(const Phys& ph, double rhoL, double rhoR, double uL, double uR, double vL, double vR,
                                                        double wL, double wR, double pL, double pR, double& rho, double& u, double& v, double& w, double& p, double& h) {
    const double betaL = 0.5 * rhoL / pL, betaR = 0.5 * rhoR / pR;
    chandrasekar_mean_state_beta(ph, rhoL, rhoR, uL, uR, vL, vR, wL, wR, betaL, betaR, rho, u, v, w, p, h);
}
__device__ __forceinline__ void chandrasekar_mean_state_beta(const Phys& ph, double rhoL, double rhoR, double uL, double uR, double vL, double vR,
                                                             double wL, double wR, double betaL, double betaR, double& rho, double& u, double& v, double& w, double& p, double& h) {
    const double betaLog = log_mean(betaL, betaR);
    rho = log_mean(rhoL, rhoR);
    u = 0.5 * (uL + uR); v = 0.5 * (vL + vR); w = 0.5 * (wL + wR);
    p = 0.5 * (rhoL + rhoR) / (betaL + betaR);
    h = 0.5 / (betaLog * ph.gm1) - 0.5 * (0.5 * ((pow2(uL) + pow2(vL) + pow2(wL)) + (pow2(uR) + pow2(vR) + pow2(wR))))
        + p / rho + pow2(u) + pow2(v) + pow2(w);
}

// Ducros, Morinishi, entropy-conserving and Chandrasekar averages (RiemannSolvers_NS.f90:1821-1886, 1964-2077).
// Out of line: these are not on the headline path and must not weigh on its register allocation.
__device__ __noinline__ void averaged_states_ext(const Phys& ph, const double QL[5], const double QR[5], double pL, double pR,
                                                 double invRhoL, double invRhoR, double flux[5]) {
    const double uL = invRhoL * QL[1], uR = invRhoR * QR[1];
    const double vL = invRhoL * QL[2], vR = invRhoR * QR[2];
    const double wL = invRhoL * QL[3], wR = invRhoR * QR[3];
    if (ph.averaging == H3D_AVG_DUCROS) {
        flux[0] = 0.25 * (QL[0] + QR[0]) * (uL + uR);
        flux[1] = 0.25 * (QL[1] + QR[1]) * (uL + uR) + 0.5 * (pL + pR);
        flux[2] = 0.25 * (QL[2] + QR[2]) * (uL + uR);
        flux[3] = 0.25 * (QL[3] + QR[3]) * (uL + uR);
        flux[4] = 0.25 * (QL[4] + pL + QR[4] + pR) * (uL + uR);
    } else if (ph.averaging == H3D_AVG_MORINISHI) {
        const double cp = ph.gamma * (1.0 / ph.gm1);
        const double hL = cp * pL, hR = cp * pR;
        flux[0] = 0.5 * (QL[1] + QR[1]);
        flux[1] = 0.25 * (QL[1] + QR[1]) * (uL + uR) + 0.5 * (pL + pR);
        flux[2] = 0.25 * (QL[1] + QR[1]) * (vL + vR);
        flux[3] = 0.25 * (QL[1] + QR[1]) * (wL + wR);
        flux[4] = 0.5 * (uL * hL + uR * hR) + 0.25 * (QL[1] * uL + QR[1] * uR) * (uL + uR)
                  + 0.25 * (QL[1] * vL + QR[1] * vR) * (vL + vR)
                  + 0.25 * (QL[1] * wL + QR[1] * wR) * (wL + wR)
                  - 0.25 * (QL[1] * pow2(uL) + QR[1] * pow2(uR))
                  - 0.25 * (QL[1] * pow2(vL) + QR[1] * pow2(vR))
                  - 0.25 * (QL[1] * pow2(wL) + QR[1] * pow2(wR));
    } else {
        double rho, u, v, w, p, h;
        if (ph.averaging == H3D_AVG_ENTROPYCONS) ec_mean_state(ph, false, QL[0], QR[0], uL, uR, vL, vR, wL, wR, pL, pR, rho, u, v, w, p, h);
        else chandrasekar_mean_state(ph, QL[0], QR[0], uL, uR, vL, vR, wL, wR, pL, pR, rho, u, v, w, p, h);
        flux[0] = rho * u; flux[1] = rho * u * u + p; flux[2] = rho * u * v; flux[3] = rho * u * w; flux[4] = rho * u * h;
    }
}

// RiemannSolvers_NS.f90:1784-1962 averaging functions on rotated states
template <bool EXT>
__device__ __forceinline__ void averaged_states(const Phys& ph, const double QL[5], const double QR[5], double pL, double pR,
                                                double invRhoL, double invRhoR, double flux[5]) {
    const double uL = invRhoL * QL[1], uR = invRhoR * QR[1];
    const double vL = invRhoL * QL[2], vR = invRhoR * QR[2];
    const double wL = invRhoL * QL[3], wR = invRhoR * QR[3];
    if constexpr (EXT) { if (ph.averaging > H3D_AVG_PIROZZOLI) { averaged_states_ext(ph, QL, QR, pL, pR, invRhoL, invRhoR, flux); return; } }
    if (ph.averaging == H3D_AVG_STANDARD) {
        flux[0] = 0.5 * (QL[1] + QR[1]);
        flux[1] = 0.5 * (QL[1] * uL + QR[1] * uR + pL + pR);
        flux[2] = 0.5 * (QL[1] * vL + QR[1] * vR);
        flux[3] = 0.5 * (QL[1] * wL + QR[1] * wR);
        flux[4] = 0.5 * (uL * (QL[4] + pL) + uR * (QR[4] + pR));
    } else {
        const double rho = 0.5 * (QL[0] + QR[0]), u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR), p = 0.5 * (pL + pR);
        flux[0] = rho * u; flux[1] = rho * u * u + p; flux[2] = rho * u * v; flux[3] = rho * u * w;
        if (ph.averaging == H3D_AVG_KENNEDYGRUBER) {
            const double e = 0.5 * (QL[4] * invRhoL + QR[4] * invRhoR);
            flux[4] = rho * u * e + p * u;
        } else {   // Pirozzoli
            const double h = 0.5 * ((QL[4] + pL) * invRhoL + (QR[4] + pR) * invRhoR);
            flux[4] = rho * u * h;
        }
    }
}

// RiemannSolvers_NS.f90:375-428 (Central) and :1251-1334 (Lax-Friedrichs); central == LxF with lambdaStab = 0
// up to the reference's own code path (no stabilisation term evaluated).
template <bool EXT>
__device__ __forceinline__ void rotated_riemann(const Phys& ph, int mode, const double QLeft[5], const double QRight[5],
                                                const double nHat[3], const double t1[3], const double t2[3], double flux[5]) {
    const double rhoL = QLeft[0], rhoR = QRight[0], invRhoL = 1.0 / rhoL, invRhoR = 1.0 / rhoR;
    const double rhouL = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    const double rhovL = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    const double rhowL = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    const double rhouR = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    const double rhovR = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    const double rhowR = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    const double rhoeL = QLeft[4], rhoeR = QRight[4];
    const double rhoV2L = (pow2(rhouL) + pow2(rhovL) + pow2(rhowL)) * invRhoL;
    const double rhoV2R = (pow2(rhouR) + pow2(rhovR) + pow2(rhowR)) * invRhoR;
    const double pL = ph.gm1 * (rhoeL - 0.5 * rhoV2L), pR = ph.gm1 * (rhoeR - 0.5 * rhoV2R);
    const double QLRot[5] = {rhoL, rhouL, rhovL, rhowL, rhoeL}, QRRot[5] = {rhoR, rhouR, rhovR, rhowR, rhoeR};
    averaged_states<EXT>(ph, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    if (mode != H3D_RIEMANN_CENTRAL) {   // Lax-Friedrichs: lambda = max(|u| + a); u-diss (:1336-1417): lambda = max(|u|)
        double lambda;
        if (mode == H3D_RIEMANN_LXF) {
            const double aL = sqrt(ph.gamma * pL * invRhoL), aR = sqrt(ph.gamma * pR * invRhoR);
            lambda = fmax(fabs(rhouL * invRhoL) + aL, fabs(rhouR * invRhoR) + aR);
        } else {
            lambda = fmax(fabs(rhouL * invRhoL), fabs(rhouR * invRhoR));
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) { const double stab = 0.5 * lambda * (QRRot[q] - QLRot[q]); flux[q] = flux[q] - ph.lambdaStab * stab; }
    }
    const double f2 = flux[1], f3 = flux[2], f4 = flux[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// RiemannSolvers_NS.f90:1541-1656 (RoeRiemannSolver): one-wave upwinding with entropy fix, uses nHat only
__device__ __forceinline__ void roe_riemann(const Phys& ph, const double QLeft[5], const double QRight[5], const double nHat[3], double flux[5]) {
    const double gamma = ph.gamma;
    const double rho = QLeft[0], rhou = QLeft[1], rhov = QLeft[2], rhow = QLeft[3], rhoe = QLeft[4];
    const double rhon = QRight[0], rhoun = QRight[1], rhovn = QRight[2], rhown = QRight[3], rhoen = QRight[4];
    const double ul = rhou / rho, vl = rhov / rho, wl = rhow / rho;
    const double pleft = (gamma - 1.0) * (rhoe - 0.5 / rho * (rhou * rhou + rhov * rhov + rhow * rhow));
    const double ur = rhoun / rhon, vr = rhovn / rhon, wr = rhown / rhon;
    const double pright = (gamma - 1.0) * (rhoen - 0.5 / rhon * (rhoun * rhoun + rhovn * rhovn + rhown * rhown));
    const double ql = nHat[0] * ul + nHat[1] * vl + nHat[2] * wl;
    const double qr = nHat[0] * ur + nHat[1] * vr + nHat[2] * wr;
    const double hl = 0.5 * (ul * ul + vl * vl + wl * wl) + gamma / (gamma - 1.0) * pleft / rho;
    const double hr = 0.5 * (ur * ur + vr * vr + wr * wr) + gamma / (gamma - 1.0) * pright / rhon;
    const double rtd = sqrt(rho * rhon);
    const double betal = rho / (rho + rtd), betar = 1.0 - betal;
    const double utd = betal * ul + betar * ur, vtd = betal * vl + betar * vr, wtd = betal * wl + betar * wr, htd = betal * hl + betar * hr;
    const double atd2 = (gamma - 1.0) * (htd - 0.5 * (utd * utd + vtd * vtd + wtd * wtd));
    const double atd = sqrt(atd2);
    const double qtd = utd * nHat[0] + vtd * nHat[1] + wtd * nHat[2];
    if (qtd >= 0.0) {
        const double dw1 = 0.5 * ((pright - pleft) / atd2 - (qr - ql) * rtd / atd);
        const double sp1 = qtd - atd;
        const double sp1m = fmin(sp1, 0.0);
        const double hd1m = ((gamma + 1.0) / 4.0 * atd / rtd) * dw1;
        const double eta1 = fmax(-fabs(sp1) - hd1m, 0.0);
        const double udw1 = dw1 * (sp1m - 0.5 * eta1);
        const double rql = rho * ql;
        flux[0] = rql + udw1;
        flux[1] = rql * ul + pleft * nHat[0] + udw1 * (utd - atd * nHat[0]);
        flux[2] = rql * vl + pleft * nHat[1] + udw1 * (vtd - atd * nHat[1]);
        flux[3] = rql * wl + pleft * nHat[2] + udw1 * (wtd - atd * nHat[2]);
        flux[4] = rql * hl + udw1 * (htd - qtd * atd);
    } else {
        const double dw4 = 0.5 * ((pright - pleft) / atd2 + (qr - ql) * rtd / atd);
        const double sp4 = qtd + atd;
        const double sp4p = fmax(sp4, 0.0);
        const double hd4 = ((gamma + 1.0) / 4.0 * atd / rtd) * dw4;
        const double eta4 = fmax(-fabs(sp4) + hd4, 0.0);
        const double udw4 = dw4 * (sp4p + 0.5 * eta4);
        const double rqr = rhon * qr;
        flux[0] = rqr - udw4;
        flux[1] = rqr * ur + pright * nHat[0] - udw4 * (utd + atd * nHat[0]);
        flux[2] = rqr * vr + pright * nHat[1] - udw4 * (vtd + atd * nHat[1]);
        flux[3] = rqr * wr + pright * nHat[2] - udw4 * (wtd + atd * nHat[2]);
        flux[4] = rqr * hr - udw4 * (htd + qtd * atd);
    }
}

// RiemannSolvers_NS.f90:1661-1762 (RusanovRiemannSolver): unrotated, smax = max(a + |q|)
__device__ __noinline__ void rusanov_riemann(const Phys& ph, const double QLeft[5], const double QRight[5], const double nHat[3], double flux[5]) {
    const double gamma = ph.gamma;
    const double rho = QLeft[0], rhou = QLeft[1], rhov = QLeft[2], rhow = QLeft[3], rhoe = QLeft[4];
    const double rhon = QRight[0], rhoun = QRight[1], rhovn = QRight[2], rhown = QRight[3], rhoen = QRight[4];
    const double ul = rhou / rho, vl = rhov / rho, wl = rhow / rho;
    const double pleft = (gamma - 1.0) * (rhoe - 0.5 / rho * (rhou * rhou + rhov * rhov + rhow * rhow));
    const double ur = rhoun / rhon, vr = rhovn / rhon, wr = rhown / rhon;
    const double pright = (gamma - 1.0) * (rhoen - 0.5 / rhon * (rhoun * rhoun + rhovn * rhovn + rhown * rhown));
    const double ql = nHat[0] * ul + nHat[1] * vl + nHat[2] * wl;
    const double qr = nHat[0] * ur + nHat[1] * vr + nHat[2] * wr;
    const double hl = 0.5 * (ul * ul + vl * vl + wl * wl) + gamma / (gamma - 1.0) * pleft / rho;
    const double hr = 0.5 * (ur * ur + vr * vr + wr * wr) + gamma / (gamma - 1.0) * pright / rhon;
    const double ar2 = (gamma - 1.0) * (hr - 0.5 * (ur * ur + vr * vr + wr * wr));
    const double al2 = (gamma - 1.0) * (hl - 0.5 * (ul * ul + vl * vl + wl * wl));
    const double ar = sqrt(ar2), al = sqrt(al2);
    const double rql = rho * ql, rqr = rhon * qr;
    flux[0] = rql + rqr;
    flux[1] = rql * ul + pleft * nHat[0] + rqr * ur + pright * nHat[0];
    flux[2] = rql * vl + pleft * nHat[1] + rqr * vr + pright * nHat[1];
    flux[3] = rql * wl + pleft * nHat[2] + rqr * wr + pright * nHat[2];
    flux[4] = rql * hl + rqr * hr;
    const double smax = fmax(ar + fabs(qr), al + fabs(ql));
#pragma unroll
    for (int q = 0; q < 5; ++q) flux[q] = (flux[q] - smax * (QRight[q] - QLeft[q])) / 2.0;
}

// RiemannSolvers_NS.f90:430-576 (StdRoeRiemannSolver): rotated, full wave decomposition, Harten / van Leer entropy fix;
// getPrimitiveVariables / getRoeVariables: VariableConversion_NS.f90:266-294, 323-365
__device__ __noinline__ void stdroe_riemann(const Phys& ph, const double QLeft[5], const double QRight[5], const double nHat[3],
                                            const double t1[3], const double t2[3], double flux[5]) {
    const double gamma = ph.gamma, gm1 = ph.gm1;
    double QLRot[5], QRRot[5];
    QLRot[0] = QLeft[0]; QRRot[0] = QRight[0];
    QLRot[1] = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    QRRot[1] = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    QLRot[2] = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    QRRot[2] = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    QLRot[3] = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    QRRot[3] = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    QLRot[4] = QLeft[4]; QRRot[4] = QRight[4];
    const double iRL = 1.0 / QLRot[0], iRR = 1.0 / QRRot[0];
    const double VLu = QLRot[1] * iRL, VLv = QLRot[2] * iRL, VLw = QLRot[3] * iRL;
    const double VRu = QRRot[1] * iRR, VRv = QRRot[2] * iRR, VRw = QRRot[3] * iRR;
    const double VLp = gm1 * (QLRot[4] - 0.5 * (VLu * QLRot[1] + VLv * QLRot[2] + VLw * QLRot[3]));
    const double VRp = gm1 * (QRRot[4] - 0.5 * (VRu * QRRot[1] + VRv * QRRot[2] + VRw * QRRot[3]));
    const double aL = sqrt(gamma * VLp * iRL), aR = sqrt(gamma * VRp * iRR);
    const double sqrtRhoL = sqrt(QLRot[0]), sqrtRhoR = sqrt(QRRot[0]);
    const double invSum = 1.0 / (sqrtRhoL + sqrtRhoR);
    const double HL = (VLp + QLRot[4]) * iRL, HR = (VRp + QRRot[4]) * iRR;
    const double u = (sqrtRhoL * VLu + sqrtRhoR * VRu) * invSum;
    const double v = (sqrtRhoL * VLv + sqrtRhoR * VRv) * invSum;
    const double w = (sqrtRhoL * VLw + sqrtRhoR * VRw) * invSum;
    const double H = (sqrtRhoL * HL + sqrtRhoR * HR) * invSum;
    const double V2 = pow2(u) + pow2(v) + pow2(w);
    const double a = sqrt(gm1 * (H - 0.5 * V2));
    double lambda[5] = {u - a, u, u, u, u + a};
    const double K[5][5] = {{1.0, u - a, v, w, H - u * a}, {1.0, u, v, w, 0.5 * V2}, {0.0, 0.0, 1.0, 0.0, v}, {0.0, 0.0, 0.0, 1.0, w}, {1.0, u + a, v, w, H + u * a}};
    double dQ[5], alpha[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) dQ[q] = QRRot[q] - QLRot[q];
    alpha[2] = dQ[2] - v * dQ[0]; alpha[3] = dQ[3] - w * dQ[0];
    dQ[4] = dQ[4] - alpha[2] * v - alpha[3] * w;
    alpha[1] = gm1 * (dQ[0] * (H - u * u) + u * dQ[1] - dQ[4]) / (pow2(a));
    alpha[0] = 0.5 * (dQ[0] * lambda[4] - dQ[1] - a * alpha[1]) / a;
    alpha[4] = dQ[0] - alpha[0] - alpha[1];
    double dLambda = fmax((VRu - aR) - (VLu - aL), 0.0);
    if (fabs(lambda[0]) >= 2.0 * dLambda) lambda[0] = fabs(lambda[0]);
    else lambda[0] = pow2(lambda[0]) / (4.0 * dLambda) + dLambda;
    dLambda = fmax((VRu + aR) - (VLu + aL), 0.0);
    if (fabs(lambda[4]) >= 2.0 * dLambda) lambda[4] = fabs(lambda[4]);
    else lambda[4] = pow2(lambda[4]) / (4.0 * dLambda) + dLambda;
    averaged_states<true>(ph, QLRot, QRRot, VLp, VRp, iRL, iRR, flux);
    if (ph.averaging == H3D_AVG_PIROZZOLI || ph.averaging == H3D_AVG_KENNEDYGRUBER) lambda[0] = lambda[4];
    double stab[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int q = 0; q < 5; ++q) stab[q] = stab[q] + 0.5 * alpha[i] * fabs(lambda[i]) * K[i][q];
#pragma unroll
    for (int q = 0; q < 5; ++q) flux[q] = flux[q] - ph.lambdaStab * stab[q];
    const double f2 = flux[1], f3 = flux[2], f4 = flux[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// shared tail of the Roe-type solvers (RiemannSolvers_NS.f90:520-572 and the identical blocks of the variants)
__device__ __forceinline__ void roe_type_tail(const Phys& ph, const double QLRot[5], const double QRRot[5], double pL, double pR, double invRhoL, double invRhoR,
                                              double uL, double uR, double aL, double aR, double u, double v, double w, double H, double a, double V2,
                                              const double alpha[5], const double nHat[3], const double t1[3], const double t2[3], double flux[5]) {
    double lambda[5] = {u - a, u, u, u, u + a};
    const double K[5][5] = {{1.0, u - a, v, w, H - u * a}, {1.0, u, v, w, 0.5 * V2}, {0.0, 0.0, 1.0, 0.0, v}, {0.0, 0.0, 0.0, 1.0, w}, {1.0, u + a, v, w, H + u * a}};
    double dLambda = fmax((uR - aR) - (uL - aL), 0.0);
    if (fabs(lambda[0]) >= 2.0 * dLambda) lambda[0] = fabs(lambda[0]);
    else lambda[0] = pow2(lambda[0]) / (4.0 * dLambda) + dLambda;
    dLambda = fmax((uR + aR) - (uL + aL), 0.0);
    if (fabs(lambda[4]) >= 2.0 * dLambda) lambda[4] = fabs(lambda[4]);
    else lambda[4] = pow2(lambda[4]) / (4.0 * dLambda) + dLambda;
    averaged_states<true>(ph, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    if (ph.averaging == H3D_AVG_PIROZZOLI || ph.averaging == H3D_AVG_KENNEDYGRUBER) lambda[0] = lambda[4];
    double stab[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int q = 0; q < 5; ++q) stab[q] = stab[q] + 0.5 * alpha[i] * fabs(lambda[i]) * K[i][q];
#pragma unroll
    for (int q = 0; q < 5; ++q) flux[q] = flux[q] - ph.lambdaStab * stab[q];
    const double f2 = flux[1], f3 = flux[2], f4 = flux[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

__device__ __forceinline__ void rotate_state(const double Q[5], const double nHat[3], const double t1[3], const double t2[3], double R[5]) {
    R[0] = Q[0];
    R[1] = Q[1] * nHat[0] + Q[2] * nHat[1] + Q[3] * nHat[2];
    R[2] = Q[1] * t1[0] + Q[2] * t1[1] + Q[3] * t1[2];
    R[3] = Q[1] * t2[0] + Q[2] * t2[1] + Q[3] * t2[2];
    R[4] = Q[4];
}

// RiemannSolvers_NS.f90:721-861 (RoePikeRiemannSolver)
__device__ __noinline__ void roepike_riemann(const Phys& ph, const double QLeft[5], const double QRight[5], const double nHat[3],
                                             const double t1[3], const double t2[3], double flux[5]) {
    const double gamma = ph.gamma, gm1 = ph.gm1;
    double QLRot[5], QRRot[5];
    rotate_state(QLeft, nHat, t1, t2, QLRot); rotate_state(QRight, nHat, t1, t2, QRRot);
    const double iRL = 1.0 / QLRot[0], iRR = 1.0 / QRRot[0];
    const double VLu = QLRot[1] * iRL, VLv = QLRot[2] * iRL, VLw = QLRot[3] * iRL;
    const double VRu = QRRot[1] * iRR, VRv = QRRot[2] * iRR, VRw = QRRot[3] * iRR;
    const double VLp = gm1 * (QLRot[4] - 0.5 * (VLu * QLRot[1] + VLv * QLRot[2] + VLw * QLRot[3]));
    const double VRp = gm1 * (QRRot[4] - 0.5 * (VRu * QRRot[1] + VRv * QRRot[2] + VRw * QRRot[3]));
    const double aL = sqrt(gamma * VLp * iRL), aR = sqrt(gamma * VRp * iRR);
    const double sqrtRhoL = sqrt(QLRot[0]), sqrtRhoR = sqrt(QRRot[0]);
    const double invSum = 1.0 / (sqrtRhoL + sqrtRhoR);
    const double HL = (VLp + QLRot[4]) * iRL, HR = (VRp + QRRot[4]) * iRR;
    const double rho = sqrtRhoL * sqrtRhoR;
    const double u = (sqrtRhoL * VLu + sqrtRhoR * VRu) * invSum;
    const double v = (sqrtRhoL * VLv + sqrtRhoR * VRv) * invSum;
    const double w = (sqrtRhoL * VLw + sqrtRhoR * VRw) * invSum;
    const double H = (sqrtRhoL * HL + sqrtRhoR * HR) * invSum;
    const double V2 = pow2(u) + pow2(v) + pow2(w);
    const double a = sqrt(gm1 * (H - 0.5 * V2));
    double alpha[5];
    alpha[0] = ((VRp - VLp) - rho * a * (VRu - VLu)) / (2.0 * a * a);
    alpha[1] = (QRight[0] - QLeft[0]) - (VRp - VLp) / (a * a);
    alpha[2] = rho * (VRv - VLv);
    alpha[3] = rho * (VRw - VLw);
    alpha[4] = ((VRp - VLp) + rho * a * (VRu - VLu)) / (2.0 * a * a);
    roe_type_tail(ph, QLRot, QRRot, VLp, VRp, iRL, iRR, VLu, VRu, aL, aR, u, v, w, H, a, V2, alpha, nHat, t1, t2, flux);
}

// RiemannSolvers_NS.f90:863-1058 (LowDissipationRoeRiemannSolver)
__device__ __noinline__ void lowdiss_riemann(const Phys& ph, const double QLeft[5], const double QRight[5], const double nHat[3],
                                             const double t1[3], const double t2[3], double flux[5]) {
    const double gamma = ph.gamma, gm1 = ph.gm1;
    double QLRot[5], QRRot[5];
    rotate_state(QLeft, nHat, t1, t2, QLRot); rotate_state(QRight, nHat, t1, t2, QRRot);
    const double rhoL = QLRot[0], rhoR = QRRot[0], invRhoL = 1.0 / rhoL, invRhoR = 1.0 / rhoR;
    const double sqrtRhoL = sqrt(rhoL), sqrtRhoR = sqrt(rhoR);
    const double invSqrtRhoL = 1.0 / sqrtRhoL, invSqrtRhoR = 1.0 / sqrtRhoR;
    const double invSum = 1.0 / (sqrtRhoL + sqrtRhoR);
    const double uL = QLRot[1] * invRhoL, uR = QRRot[1] * invRhoR, vL = QLRot[2] * invRhoL, vR = QRRot[2] * invRhoR, wL = QLRot[3] * invRhoL, wR = QRRot[3] * invRhoR;
    const double rhoV2L = (pow2(uL) + pow2(vL) + pow2(wL)) * rhoL, rhoV2R = (pow2(uR) + pow2(vR) + pow2(wR)) * rhoR;
    const double rhoHL = gamma * QLRot[4] - 0.5 * gm1 * rhoV2L, rhoHR = gamma * QRRot[4] - 0.5 * gm1 * rhoV2R;
    const double pL = gm1 * (QLRot[4] - 0.5 * rhoV2L), pR = gm1 * (QRRot[4] - 0.5 * rhoV2R);
    const double aL = sqrt(gamma * pL * invRhoL), aR = sqrt(gamma * pR * invRhoR);
    const double rho = sqrtRhoL * sqrtRhoR;
    const double u = (invSqrtRhoL * QLRot[1] + invSqrtRhoR * QRRot[1]) * invSum;
    const double v = (invSqrtRhoL * QLRot[2] + invSqrtRhoR * QRRot[2]) * invSum;
    const double w = (invSqrtRhoL * QLRot[3] + invSqrtRhoR * QRRot[3]) * invSum;
    const double H = (invSqrtRhoL * rhoHL + invSqrtRhoR * rhoHR) * invSum;
    const double V2abs = pow2(u) + pow2(v) + pow2(w);
    const double a = sqrt(gm1 * (H - 0.5 * V2abs));
    const double ML = fabs(uL) / aL, MR = fabs(uR) / aR;
    const double z = fmin(1.0, fmax(ML, MR));
    const double du = z * (uR - uL), dv = z * (vR - vL), dw = z * (wR - wL), dp = pR - pL;
    double alpha[5];
    alpha[0] = (dp - rho * a * du) / (2.0 * a * a);
    alpha[1] = (rhoR - rhoL) - dp / (a * a);
    alpha[2] = rho * dv;
    alpha[3] = rho * dw;
    alpha[4] = (dp + rho * a * du) / (2.0 * a * a);
    roe_type_tail(ph, QLRot, QRRot, pL, pR, invRhoL, invRhoR, uL, uR, aL, aR, u, v, w, H, a, V2abs, alpha, nHat, t1, t2, flux);
}

// RiemannSolvers_NS.f90:578-719 (MatrixDissipationRiemannSolver); entropy variables: VariableConversion_NS.f90:211-237
__device__ __noinline__ void matrixdiss_riemann(const Phys& ph, const double QLeft[5], const double QRight[5], const double nHat[3],
                                                const double t1[3], const double t2[3], double flux[5]) {
    const double gamma = ph.gamma, gm1 = ph.gm1;
    const double invGamma = 1.0 / gamma, cp = gamma / gm1, gammaMinus1Div2g = gm1 / (2.0 * gamma), invGammaMinus1 = 1.0 / gm1;
    double QLRot[5], QRRot[5], EVL[5], EVR[5];
    rotate_state(QLeft, nHat, t1, t2, QLRot); rotate_state(QRight, nHat, t1, t2, QRRot);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const double* Q = side ? QRRot : QLRot; double* U = side ? EVR : EVL;
        const double invRho = 1.0 / Q[0];
        const double rhoV2 = (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) * invRho;
        const double p = gm1 * (Q[4] - 0.5 * rhoV2);
        const double invP = 1.0 / p;
        U[0] = (gamma - (log(p) - gamma * log(Q[0]))) * invGammaMinus1 - 0.5 * rhoV2 * invP;
        U[1] = Q[1] * invP; U[2] = Q[2] * invP; U[3] = Q[3] * invP; U[4] = -Q[0] * invP;
    }
    const double invRhoL = 1.0 / QLRot[0], invRhoR = 1.0 / QRRot[0];
    const double uL = QLRot[1] * invRhoL, uR = QRRot[1] * invRhoR, vL = QLRot[2] * invRhoL, vR = QRRot[2] * invRhoR, wL = QLRot[3] * invRhoL, wR = QRRot[3] * invRhoR;
    const double vtotL = uL * uL + vL * vL + wL * wL, vtotR = uR * uR + vR * vR + wR * wR;
    const double pL = gm1 * (QLRot[4] - 0.5 * QLRot[0] * vtotL), pR = gm1 * (QRRot[4] - 0.5 * QRRot[0] * vtotR);
    const double betaL = -0.5 * EVL[4], betaR = -0.5 * EVR[4];
    const double betaLogMean = log_mean(betaL, betaR), rhoLogMean = log_mean(QLRot[0], QRRot[0]);
    const double pMean = 0.5 * (QLRot[0] + QRRot[0]) / (betaL + betaR);
    const double a_bar = sqrt(gamma * pMean / rhoLogMean);
    const double uMean = 0.5 * (uL + uR), vMean = 0.5 * (vL + vR), wMean = 0.5 * (wL + wR);
    const double V2abs = 2.0 * (pow2(uMean) + pow2(vMean) + pow2(wMean)) - 0.5 * (vtotL + vtotR);
    const double h_bar = 0.5 * (cp / betaLogMean + V2abs);
    double lambda[5] = {fabs(uMean - a_bar), fabs(uMean), fabs(uMean), fabs(uMean), fabs(uMean + a_bar)};
    const double R1[5][5] = {{1.0, 1.0, 0.0, 0.0, 1.0},
                             {uMean - a_bar, uMean, 0.0, 0.0, uMean + a_bar},
                             {vMean, vMean, 1.0, 0.0, vMean},
                             {wMean, wMean, 0.0, 1.0, wMean},
                             {h_bar - uMean * a_bar, 0.5 * V2abs, vMean, wMean, h_bar + uMean * a_bar}};
    double T[5];
    T[0] = 0.5 * rhoLogMean * invGamma; T[1] = 2.0 * gammaMinus1Div2g * rhoLogMean; T[2] = pMean; T[3] = pMean; T[4] = T[0];
    averaged_states<true>(ph, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    if (ph.averaging == H3D_AVG_PIROZZOLI || ph.averaging == H3D_AVG_KENNEDYGRUBER) lambda[0] = lambda[4];
    double stab[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < 5; ++k) stab[i] = stab[i] + 0.5 * R1[i][j] * lambda[j] * T[j] * R1[k][j] * (EVR[k] - EVL[k]);
#pragma unroll
    for (int q = 0; q < 5; ++q) flux[q] = flux[q] - ph.lambdaStab * stab[q];
    const double f2 = flux[1], f3 = flux[2], f4 = flux[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// EXT = false: Roe, Lax-Friedrichs, central with the standard / Kennedy-Gruber / Pirozzoli averages (the instantiation on
// the headline path); EXT = true adds the out-of-line solvers and averages
template <bool EXT>
__device__ __forceinline__ void riemann_solver(const Phys& ph, const double QL[5], const double QR[5], const double nHat[3],
                                               const double t1[3], const double t2[3], double flux[5]) {
    if (ph.riemann == H3D_RIEMANN_ROE) { roe_riemann(ph, QL, QR, nHat, flux); return; }
    if constexpr (EXT) {
        if (ph.riemann == H3D_RIEMANN_RUSANOV) { rusanov_riemann(ph, QL, QR, nHat, flux); return; }
        if (ph.riemann == H3D_RIEMANN_STDROE) { stdroe_riemann(ph, QL, QR, nHat, t1, t2, flux); return; }
        if (ph.riemann == H3D_RIEMANN_ROEPIKE) { roepike_riemann(ph, QL, QR, nHat, t1, t2, flux); return; }
        if (ph.riemann == H3D_RIEMANN_LOWDISSROE) { lowdiss_riemann(ph, QL, QR, nHat, t1, t2, flux); return; }
        if (ph.riemann == H3D_RIEMANN_MATRIXDISS) { matrixdiss_riemann(ph, QL, QR, nHat, t1, t2, flux); return; }
    }
    rotated_riemann<EXT>(ph, ph.riemann, QL, QR, nHat, t1, t2, flux);
}

// Morinishi, Ducros, entropy-conserving and Chandrasekar two-point fluxes (RiemannSolvers_NS.f90:2147-2294, 2439-2641)
__device__ __noinline__ void two_point_flux_ext(const Phys& ph, const double QL[5], const double QR[5], const double JaL[3], const double JaR[3], double fs[5]) {
    const double invRhoL = 1.0 / QL[0], invRhoR = 1.0 / QR[0];
    const double uL = invRhoL * QL[1], uR = invRhoR * QR[1];
    const double vL = invRhoL * QL[2], vR = invRhoR * QR[2];
    const double wL = invRhoL * QL[3], wR = invRhoR * QR[3];
    const double pL = ph.gm1 * (QL[4] - 0.5 * (QL[1] * uL + QL[2] * vL + QL[3] * wL));
    const double pR = ph.gm1 * (QR[4] - 0.5 * (QR[1] * uR + QR[2] * vR + QR[3] * wR));
    const double Ja[3] = {0.5 * (JaL[0] + JaR[0]), 0.5 * (JaL[1] + JaR[1]), 0.5 * (JaL[2] + JaR[2])};
    double F[3][5];
    if (ph.averaging == H3D_AVG_DUCROS) {
        const double velSum[3] = {uL + uR, vL + vR, wL + wR};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            F[d][0] = 0.25 * (QL[0] + QR[0]) * velSum[d];
#pragma unroll
            for (int c = 0; c < 3; ++c) F[d][1 + c] = 0.25 * (QL[1 + c] + QR[1 + c]) * velSum[d];
            F[d][1 + d] = 0.25 * (QL[1 + d] + QR[1 + d]) * velSum[d] + 0.5 * (pL + pR);
            F[d][4] = 0.25 * (QL[4] + pL + QR[4] + pR) * velSum[d];
        }
    } else if (ph.averaging == H3D_AVG_MORINISHI) {
        const double cp = ph.gamma * (1.0 / ph.gm1);
        const double hL = cp * pL, hR = cp * pR;
        const double velL[3] = {uL, vL, wL}, velR[3] = {uR, vR, wR};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double mL = QL[1 + d], mR = QR[1 + d];
            F[d][0] = 0.5 * (mL + mR);
#pragma unroll
            for (int c = 0; c < 3; ++c) F[d][1 + c] = 0.25 * (mL + mR) * (velL[c] + velR[c]);
            F[d][1 + d] = 0.25 * (mL + mR) * (velL[d] + velR[d]) + 0.5 * (pL + pR);
            F[d][4] = 0.5 * (velL[d] * hL + velR[d] * hR) + 0.25 * (mL * uL + mR * uR) * (uL + uR)
                      + 0.25 * (mL * vL + mR * vR) * (vL + vR)
                      + 0.25 * (mL * wL + mR * wR) * (wL + wR)
                      - 0.25 * (mL * pow2(uL) + mR * pow2(uR))
                      - 0.25 * (mL * pow2(vL) + mR * pow2(vR))
                      - 0.25 * (mL * pow2(wL) + mR * pow2(wR));
        }
    } else {
        double rho, u, v, w, p, h;
        if (ph.averaging == H3D_AVG_ENTROPYCONS) ec_mean_state(ph, true, QL[0], QR[0], uL, uR, vL, vR, wL, wR, pL, pR, rho, u, v, w, p, h);
        else chandrasekar_mean_state(ph, QL[0], QR[0], uL, uR, vL, vR, wL, wR, pL, pR, rho, u, v, w, p, h);
        F[0][0] = rho * u; F[0][1] = rho * u * u + p; F[0][2] = rho * u * v; F[0][3] = rho * u * w; F[0][4] = rho * u * h;
        F[1][0] = rho * v; F[1][1] = rho * v * u; F[1][2] = rho * v * v + p; F[1][3] = rho * v * w; F[1][4] = rho * v * h;
        F[2][0] = rho * w; F[2][1] = rho * w * u; F[2][2] = rho * w * v; F[2][3] = rho * w * w + p; F[2][4] = rho * w * h;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) fs[q] = F[0][q] * Ja[0] + F[1][q] * Ja[1] + F[2][q] * Ja[2];
}

// RiemannSolvers_NS.f90:2085-2145 (StandardDG), :2296-2367 (Kennedy-Gruber), :2369-2437 (Pirozzoli) two-point fluxes
template <bool EXT>
__device__ __forceinline__ void two_point_flux(const Phys& ph, const double QL[5], const double QR[5], const double JaL[3], const double JaR[3], double fs[5]) {
    const double invRhoL = 1.0 / QL[0], invRhoR = 1.0 / QR[0];
    const double uL = invRhoL * QL[1], uR = invRhoR * QR[1];
    const double vL = invRhoL * QL[2], vR = invRhoR * QR[2];
    const double wL = invRhoL * QL[3], wR = invRhoR * QR[3];
    const double pL = ph.gm1 * (QL[4] - 0.5 * (QL[1] * uL + QL[2] * vL + QL[3] * wL));
    const double pR = ph.gm1 * (QR[4] - 0.5 * (QR[1] * uR + QR[2] * vR + QR[3] * wR));
    double f[5], g[5], h[5];
    if constexpr (EXT) { if (ph.averaging > H3D_AVG_PIROZZOLI) { two_point_flux_ext(ph, QL, QR, JaL, JaR, fs); return; } }
    if (ph.averaging == H3D_AVG_STANDARD) {
        const double Ja[3] = {JaL[0] + JaR[0], JaL[1] + JaR[1], JaL[2] + JaR[2]};
        f[0] = QL[1] + QR[1]; f[1] = QL[1] * uL + QR[1] * uR + pL + pR; f[2] = QL[1] * vL + QR[1] * vR; f[3] = QL[1] * wL + QR[1] * wR;
        f[4] = uL * (QL[4] + pL) + uR * (QR[4] + pR);
        g[0] = QL[2] + QR[2]; g[1] = QL[2] * uL + QR[2] * uR; g[2] = QL[2] * vL + QR[2] * vR + pL + pR; g[3] = QL[2] * wL + QR[2] * wR;
        g[4] = vL * (QL[4] + pL) + vR * (QR[4] + pR);
        h[0] = QL[3] + QR[3]; h[1] = QL[3] * uL + QR[3] * uR; h[2] = QL[3] * vL + QR[3] * vR; h[3] = QL[3] * wL + QR[3] * wR + pL + pR;
        h[4] = wL * (QL[4] + pL) + wR * (QR[4] + pR);
#pragma unroll
        for (int q = 0; q < 5; ++q) fs[q] = 0.25 * (f[q] * Ja[0] + g[q] * Ja[1] + h[q] * Ja[2]);
        return;
    }
    const double rho = 0.5 * (QL[0] + QR[0]), u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR), p = 0.5 * (pL + pR);
    const double Ja[3] = {0.5 * (JaL[0] + JaR[0]), 0.5 * (JaL[1] + JaR[1]), 0.5 * (JaL[2] + JaR[2])};
    f[0] = rho * u; f[1] = rho * u * u + p; f[2] = rho * u * v; f[3] = rho * u * w;
    g[0] = rho * v; g[1] = rho * v * u; g[2] = rho * v * v + p; g[3] = rho * v * w;
    h[0] = rho * w; h[1] = rho * w * u; h[2] = rho * w * v; h[3] = rho * w * w + p;
    if (ph.averaging == H3D_AVG_KENNEDYGRUBER) {
        const double e = 0.5 * (QL[4] * invRhoL + QR[4] * invRhoR);
        f[4] = rho * u * e + p * u; g[4] = rho * v * e + p * v; h[4] = rho * w * e + p * w;
    } else {
        const double hh = 0.5 * ((QL[4] + pL) * invRhoL + (QR[4] + pR) * invRhoR);
        f[4] = rho * u * hh; g[4] = rho * v * hh; h[4] = rho * w * hh;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) fs[q] = f[q] * Ja[0] + g[q] * Ja[1] + h[q] * Ja[2];
}

// The same two-point fluxes from per-node primitives P = [rho, u, v, w, p, X] evaluated ONCE per node instead of once per
// pair (the split-form volume term evaluates 3 N pairs per node): X = rho e / rho for Kennedy-Gruber, (rho e + p) / rho for
// Pirozzoli, beta = 1/2 rho / p for Chandrasekar, unused by the entropy-conserving flux.  Every per-node expression is the one the pairwise
// routines above evaluate (invRho = 1 / rho; u = invRho * rho u; p = (gamma-1)(rho e - (rho u u + rho v v + rho w w)/2)), so the
// result is bit-identical.  prim_two_point_ok() tells which averages take this path.
__device__ __forceinline__ bool prim_two_point_ok(int averaging) {
    return averaging == H3D_AVG_KENNEDYGRUBER || averaging == H3D_AVG_PIROZZOLI || averaging == H3D_AVG_ENTROPYCONS || averaging == H3D_AVG_CHANDRASEKAR;
}
__device__ __forceinline__ void node_primitives(const Phys& ph, const double Q[5], double P[6]) {
    const double invRho = 1.0 / Q[0];
    const double u = invRho * Q[1], v = invRho * Q[2], w = invRho * Q[3];
    const double p = ph.gm1 * (Q[4] - 0.5 * (Q[1] * u + Q[2] * v + Q[3] * w));
    P[0] = Q[0]; P[1] = u; P[2] = v; P[3] = w; P[4] = p;
    P[5] = ph.averaging == H3D_AVG_KENNEDYGRUBER ? Q[4] * invRho : (ph.averaging == H3D_AVG_CHANDRASEKAR ? 0.5 * Q[0] / p : (Q[4] + p) * invRho);
}
template <bool EXT>
__device__ __forceinline__ void two_point_flux_prim(const Phys& ph, const double PL[6], const double PR[6], const double JaL[3], const double JaR[3], double fs[5]) {
    const double Ja[3] = {0.5 * (JaL[0] + JaR[0]), 0.5 * (JaL[1] + JaR[1]), 0.5 * (JaL[2] + JaR[2])};
    double f[5], g[5], h[5];
    if constexpr (EXT) {
        if (ph.averaging > H3D_AVG_PIROZZOLI) {
            double rho, u, v, w, p, hh;
            if (ph.averaging == H3D_AVG_ENTROPYCONS) ec_mean_state(ph, true, PL[0], PR[0], PL[1], PR[1], PL[2], PR[2], PL[3], PR[3], PL[4], PR[4], rho, u, v, w, p, hh);
            else chandrasekar_mean_state_beta(ph, PL[0], PR[0], PL[1], PR[1], PL[2], PR[2], PL[3], PR[3], PL[5], PR[5], rho, u, v, w, p, hh);
            f[0] = rho * u; f[1] = rho * u * u + p; f[2] = rho * u * v; f[3] = rho * u * w; f[4] = rho * u * hh;
            g[0] = rho * v; g[1] = rho * v * u; g[2] = rho * v * v + p; g[3] = rho * v * w; g[4] = rho * v * hh;
            h[0] = rho * w; h[1] = rho * w * u; h[2] = rho * w * v; h[3] = rho * w * w + p; h[4] = rho * w * hh;
#pragma unroll
            for (int q = 0; q < 5; ++q) fs[q] = f[q] * Ja[0] + g[q] * Ja[1] + h[q] * Ja[2];
            return;
        }
    }
    const double rho = 0.5 * (PL[0] + PR[0]), u = 0.5 * (PL[1] + PR[1]), v = 0.5 * (PL[2] + PR[2]), w = 0.5 * (PL[3] + PR[3]), p = 0.5 * (PL[4] + PR[4]);
    f[0] = rho * u; f[1] = rho * u * u + p; f[2] = rho * u * v; f[3] = rho * u * w;
    g[0] = rho * v; g[1] = rho * v * u; g[2] = rho * v * v + p; g[3] = rho * v * w;
    h[0] = rho * w; h[1] = rho * w * u; h[2] = rho * w * v; h[3] = rho * w * w + p;
    const double X = 0.5 * (PL[5] + PR[5]);
    if (ph.averaging == H3D_AVG_KENNEDYGRUBER) { f[4] = rho * u * X + p * u; g[4] = rho * v * X + p * v; h[4] = rho * w * X + p * w; }
    else { f[4] = rho * u * X; g[4] = rho * v * X; h[4] = rho * w * X; }
#pragma unroll
    for (int q = 0; q < 5; ++q) fs[q] = f[q] * Ja[0] + g[q] * Ja[1] + h[q] * Ja[2];
}

// Kennedy-Gruber / Pirozzoli from HALVED per-node primitives and metrics: 1/2 (a + b) = a/2 + b/2 exactly in binary floating
// point (scaling by a power of two commutes with rounding), so the nine multiplications by 1/2 of every pair are done once per node.
__device__ __forceinline__ void two_point_flux_half(const Phys& ph, const double hL[6], const double hR[6], const double hJaL[3], const double hJaR[3], double fs[5]) {
    const double Ja[3] = {hJaL[0] + hJaR[0], hJaL[1] + hJaR[1], hJaL[2] + hJaR[2]};
    const double rho = hL[0] + hR[0], u = hL[1] + hR[1], v = hL[2] + hR[2], w = hL[3] + hR[3], p = hL[4] + hR[4], X = hL[5] + hR[5];
    double f[5], g[5], h[5];
    f[0] = rho * u; f[1] = rho * u * u + p; f[2] = rho * u * v; f[3] = rho * u * w;
    g[0] = rho * v; g[1] = rho * v * u; g[2] = rho * v * v + p; g[3] = rho * v * w;
    h[0] = rho * w; h[1] = rho * w * u; h[2] = rho * w * v; h[3] = rho * w * w + p;
    if (ph.averaging == H3D_AVG_KENNEDYGRUBER) { f[4] = rho * u * X + p * u; g[4] = rho * v * X + p * v; h[4] = rho * w * X + p * w; }
    else { f[4] = rho * u * X; g[4] = rho * v * X; h[4] = rho * w * X; }
#pragma unroll
    for (int q = 0; q < 5; ++q) fs[q] = f[q] * Ja[0] + g[q] * Ja[1] + h[q] * Ja[2];
}

// ---- boundary conditions (libs/physics/common/{NoSlipWall,FreeSlipWall,Inflow,Outflow}BC.f90) -------------
// Zone parameters P[16]: walls: P[0..2] vWall, P[3] wallType (0 adiabatic / 1 isothermal), P[4] Twall,
// P[5] T_ref*gammaM2*(gamma-1) (no-slip) or T_ref*gammaM2 (free-slip), P[6] eWall; inflow: rho,u,v,w,p; outflow: P[4] pExt.
// External state for the Riemann solver (FlowState).
__device__ __forceinline__ void bc_flow_state(const Phys& ph, int type, const double* P, const double nHat[3], double Q[5]) {
    if (type == H3D_BC_NOSLIPWALL) {            // NoSlipWallBC.f90:265-296
        Q[1] = 2.0 * Q[0] * P[0] - Q[1]; Q[2] = 2.0 * Q[0] * P[1] - Q[2]; Q[3] = 2.0 * Q[0] * P[2] - Q[3];
        Q[4] = Q[4] + P[3] * (Q[0] * P[4] / P[5] - Q[4]);
    } else if (type == H3D_BC_FREESLIPWALL) {   // FreeSlipWallBC.f90:249-283
        const double qNorm = nHat[0] * Q[1] + nHat[1] * Q[2] + nHat[2] * Q[3];
        Q[1] = Q[1] - 2.0 * qNorm * nHat[0]; Q[2] = Q[2] - 2.0 * qNorm * nHat[1]; Q[3] = Q[3] - 2.0 * qNorm * nHat[2];
        const double paux = Q[0] * P[4] / P[5];
        Q[4] = Q[4] + P[3] * (paux / ph.gm1 + 0.5 * (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) / Q[0] - Q[4]);
    } else if (type == H3D_BC_INFLOW) {         // InflowBC.f90:363-406 with zero turbulence intensity
        const double u = P[1], v = P[2], w = P[3];
        Q[0] = P[0]; Q[1] = Q[0] * u; Q[2] = Q[0] * v; Q[3] = Q[0] * w;
        Q[4] = P[4] / (ph.gamma - 1.0) + 0.5 * Q[0] * (u * u + v * v + w * w);
    } else if (type == H3D_BC_OUTFLOW) {        // OutflowBC.f90:226-288
        const double pExt = P[4];
        double qDotN = (nHat[0] * Q[1] + nHat[1] * Q[2] + nHat[2] * Q[3]) / Q[0];
        const double qTanx = Q[1] / Q[0] - qDotN * nHat[0], qTany = Q[2] / Q[0] - qDotN * nHat[1], qTanz = Q[3] / Q[0] - qDotN * nHat[2];
        const double p = ph.gm1 * (Q[4] - 0.5 * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]) / Q[0]);
        const double a2 = ph.gamma * p / Q[0];
        double a = sqrt(a2);
        if (fabs(qDotN / a) <= 1.0) {
            const double rPlus = qDotN + 2.0 * a / ph.gm1;
            const double entropyConstant = p - a2 * Q[0];
            const double rho = -(entropyConstant - pExt) / a2;
            a = sqrt(ph.gamma * pExt / rho);
            qDotN = rPlus - 2.0 * a / ph.gm1;
            const double u = qTanx + qDotN * nHat[0], v = qTany + qDotN * nHat[1], w = qTanz + qDotN * nHat[2];
            Q[0] = rho; Q[1] = rho * u; Q[2] = rho * v; Q[3] = rho * w;
            Q[4] = pExt / ph.gm1 + 0.5 * rho * (u * u + v * v + w * w);
        }
    }
}

// Boundary value of the gradient variables (FlowGradVars with STATE variables); us enters as the interior state
// u_int = GetGradients(Q_int) on entry in us (BR1_ComputeBoundaryFlux, EllipticBR1.f90:703-722); GV = false: State variables
template <bool GV = false>
__device__ __forceinline__ void bc_grad_vars(const Phys& ph, int type, const double* P, const double nHat[3], const double Qi[5], double us[5]) {
    double Qa[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) Qa[q] = Qi[q];
    if (type == H3D_BC_NOSLIPWALL) {            // NoSlipWallBC.f90:298-337: U(IRHO) keeps the interior value
        const double invRho = 1.0 / Qi[0];
        const double e_int = invRho * (Qi[4] - 0.5 * invRho * (pow2(Qi[1]) + pow2(Qi[2]) + pow2(Qi[3])));
        const double U1 = us[0];
        Qa[1] = Qi[0] * P[0]; Qa[2] = Qi[0] * P[1]; Qa[3] = Qi[0] * P[2];
        Qa[4] = Qi[0] * ((1.0 - P[3]) * e_int + P[3] * P[6] + 0.5 * (P[0] * P[0] + P[1] * P[1] + P[2] * P[2]));
        if (GV) get_gradients(ph, Qa, us);
        else {
#pragma unroll
            for (int q = 0; q < 5; ++q) us[q] = Qa[q];
        }
        us[0] = U1;
    } else if (type == H3D_BC_FREESLIPWALL) {   // FreeSlipWallBC.f90:285-312
        Qa[4] = Qi[4] + P[3] * (Qi[0] * P[6] + 0.5 * (pow2(Qi[1]) + pow2(Qi[2]) + pow2(Qi[3])) / Qi[0] - Qi[4]);
        if (GV) get_gradients(ph, Qa, us);
        else {
#pragma unroll
            for (int q = 0; q < 5; ++q) us[q] = Qa[q];
        }
    } else if (type == H3D_BC_INFLOW || type == H3D_BC_OUTFLOW) {   // GenericBC_FlowGradVars (GenericBoundaryConditionClass.f90:243-278)
        double Ua[5];
        bc_flow_state(ph, type, P, nHat, Qa);
        if (GV) get_gradients(ph, Qa, Ua);
        else {
#pragma unroll
            for (int q = 0; q < 5; ++q) Ua[q] = Qa[q];
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) us[q] = 0.5 * (Ua[q] + us[q]);
    }
}

// Neumann fix-up of the boundary viscous flux (FlowNeumann); Q is the interior trace
__device__ __forceinline__ void bc_neumann(int type, const double* P, const double Q[5], double visc[5]) {
    if (type == H3D_BC_NOSLIPWALL) {            // NoSlipWallBC.f90:339-372
        const double invRho = 1.0 / Q[0], u = invRho * Q[1], v = invRho * Q[2], w = invRho * Q[3];
        const double viscWork = u * visc[1] + v * visc[2] + w * visc[3];
        const double heatFlux = visc[4] - viscWork;
        visc[0] = 0.0;
        visc[4] = (P[0] * visc[1] + P[1] * visc[2] + P[2] * visc[3]) + P[3] * heatFlux;
    } else if (type == H3D_BC_FREESLIPWALL) {   // FreeSlipWallBC.f90:314-351
        const double viscWork = (visc[1] * Q[1] + visc[2] * Q[2] + visc[3] * Q[3]) / Q[0];
        const double heatFlux = visc[4] - viscWork;
        visc[0] = 0.0; visc[1] = 0.0; visc[2] = 0.0; visc[3] = 0.0;
        visc[4] = P[3] * heatFlux;
    } else if (type == H3D_BC_INFLOW || type == H3D_BC_OUTFLOW) {
#pragma unroll
        for (int q = 0; q < 5; ++q) visc[q] = 0.0;
    }
}

}  // namespace h3d
