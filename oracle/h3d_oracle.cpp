// ======================================================================================================
//  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//  CPU restatement ("oracle") of HORSES3D's explicit compressible Navier-Stokes residual and RK step.
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
//  It restates the reference algorithm loop by loop, in the reference's own accumulation order, in plain
//  C++ compiled with -ffp-contract=off (the reference's gfortran RELEASE build emits no FMA).  Every
//  routine cites the Fortran it follows (paths relative to /root/reference/Solver/src).
//
//  Parity pins (tests/test_oracle_pins.py): this restatement reproduces, at the reference's own
//  tolerances, the regression values of its test cases Components/NodalStorage (K6), NavierStokes/TaylorGreen (K1),
//  Euler/TaylorGreenKEPEC (K2), NavierStokes/Convergence (K3, P=7), Euler/BoxAroundCircle and BoxAroundCirclePirozzoli (K4b, K4, 1000 steps; the
//  force monitor of the latter to 5e-10 where the reference asserts 1e-10), Euler/UniformFlow (K12, iterations to tolerance), NavierStokes/Cylinder, CylinderSmagorinsky, CylinderWALE and CylinderVreman (K5, K5b, K5e),
//  CylinderDucros and CylinderChandrasekarRoe (K5c, K5d), CylinderBR2 and CylinderIP (K7, K8), TaylorGreenKEP_BR2 and
//  TaylorGreenKEPEC_IP (K9), Convergence_energy and Convergence_entropy (K10, P=7), EnergyConservingTest and
//  EntropyConservingTest (K11), NavierStokes/CylinderDifferentOrders (K13: element-wise anisotropic orders, h3d_oracle_p.inc).  The reference itself cannot be built in this container (no Fortran compiler), so there is
//  no oracle/_ref; parity with the reference rests on those pins.
// ======================================================================================================
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/h3d_gpu.h"

namespace {

constexpr int NCONS = 5, NGRAD = 5, NDIM = 3;
constexpr int IRHO = 0, IRHOU = 1, IRHOV = 2, IRHOW = 3, IRHOE = 4;
constexpr int IX = 0, IY = 1, IZ = 2;
enum { EFRONT = 0, EBACK = 1, EBOTTOM = 2, ERIGHT = 3, ETOP = 4, ELEFT = 5 };
inline double POW2(double x) { return x * x; }

// ---- MeshTypes.f90:70-108
inline void leftIndexes2Right(int i, int j, int Nx, int Ny, int rot, int& ii, int& jj) {
    switch (rot) {
        case 0: ii = i; jj = j; break;
        case 1: ii = Ny - j; jj = i; break;
        case 2: ii = Nx - i; jj = Ny - j; break;
        case 3: ii = j; jj = Nx - i; break;
        case 4: ii = j; jj = i; break;
        case 5: ii = Nx - i; jj = j; break;
        case 6: ii = Ny - j; jj = Nx - i; break;
        default: ii = i; jj = Ny - j; break;
    }
}

struct Oracle {
    std::string err;
    H3dPhysics ph{};
    // basis
    int N = -1, n = 0, nodeType = H3D_GAUSS;
    std::vector<double> x, w, D, hatD, sharpD, v, b;
    // mesh
    int nElem = 0, nFace = 0;
    std::vector<int> elemFace, elemFaceSide, faceElem, faceElemSide, faceRot, faceType, faceZone;
    std::vector<double> JaXi, JaEta, JaZeta, jac, invJac, xyz, volume;
    std::vector<double> fNormal, fT1, fT2, fJac, fX, fSurface;
    std::vector<double> dWall, fdWall;      // e % geom % dWall(i,j,k), f % geom % dWall(i,j)
    std::vector<double> fH;                 // f % geom % h
    bool limited = false; double limiterMin = 1e-13;   // LIMITED, LIMITER_MIN (ExplicitMethods.f90:28-29)
    std::vector<double> stats; int statSamples = 0;
    std::vector<double> snapshot;   // e % storage % stats % data(var,i,j,k)
    int nZones = 0; std::vector<int> bcType; std::vector<double> bcParams;
    // element storage (reference order [e][k][j][i][eq])
    std::vector<double> Q, QDot, G, S, Ux, Uy, Uz, mu;   // mu: [e][node][2] = (mu, kappa)
    bool hasSource = false;
    // face storage [f][side][j][i][...]
    std::vector<double> fQ, fUx, fUy, fUz, fStar, fmu, unStar;  // unStar: [f][side][j][i][3][5]
    int n2() const { return n * n; }
    int n3() const { return n * n * n; }
    // p-nonconforming meshes (h3d_oracle_p.inc): per-element orders; the element fields above keep the reference's packed order
    bool mixed = false; void* pdata = nullptr;
};

// ------------------------------------------------------------------------------------------------
//  Physics (libs/physics/navierstokes)
// ------------------------------------------------------------------------------------------------
// Physics_NS.f90:51-97
inline void EulerFlux(const Oracle& o, const double* Q, double F[NCONS][NDIM]) {
    const double gm1 = o.ph.gammaMinus1;
    double u = Q[IRHOU] / Q[IRHO], v = Q[IRHOV] / Q[IRHO], w = Q[IRHOW] / Q[IRHO];
    double p = gm1 * (Q[IRHOE] - 0.5 * (Q[IRHOU] * u + Q[IRHOV] * v + Q[IRHOW] * w));
    F[IRHO][IX] = Q[IRHOU]; F[IRHOU][IX] = Q[IRHOU] * u + p; F[IRHOV][IX] = Q[IRHOU] * v; F[IRHOW][IX] = Q[IRHOU] * w; F[IRHOE][IX] = (Q[IRHOE] + p) * u;
    F[IRHO][IY] = Q[IRHOV]; F[IRHOU][IY] = F[IRHOV][IX]; F[IRHOV][IY] = Q[IRHOV] * v + p; F[IRHOW][IY] = Q[IRHOV] * w; F[IRHOE][IY] = (Q[IRHOE] + p) * v;
    F[IRHO][IZ] = Q[IRHOW]; F[IRHOU][IZ] = F[IRHOW][IX]; F[IRHOV][IZ] = F[IRHOW][IY]; F[IRHOW][IZ] = Q[IRHOW] * w + p; F[IRHOE][IZ] = (Q[IRHOE] + p) * w;
}

// Physics_NS.f90:246-304
inline void ViscousFlux_STATE(const Oracle& o, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z,
                              double mu, double beta, double kappa, double F[NCONS][NDIM]) {
    const double gm1 = o.ph.gammaMinus1, gM2 = o.ph.gammaM2;
    double invRho = 1.0 / Q[IRHO];
    double u = Q[IRHOU] * invRho, v = Q[IRHOV] * invRho, w = Q[IRHOW] * invRho;
    double uDivRho[3] = {u * invRho, v * invRho, w * invRho};
    double u_x[3], u_y[3], u_z[3], nablaT[3];
    for (int c = 0; c < 3; ++c) {
        u_x[c] = invRho * Q_x[IRHOU + c] - uDivRho[c] * Q_x[IRHO];
        u_y[c] = invRho * Q_y[IRHOU + c] - uDivRho[c] * Q_y[IRHO];
        u_z[c] = invRho * Q_z[IRHOU + c] - uDivRho[c] * Q_z[IRHO];
    }
    nablaT[IX] = gm1 * gM2 * (invRho * Q_x[IRHOE] - Q[IRHOE] * invRho * invRho * Q_x[IRHO] - u * u_x[IX] - v * u_x[IY] - w * u_x[IZ]);
    nablaT[IY] = gm1 * gM2 * (invRho * Q_y[IRHOE] - Q[IRHOE] * invRho * invRho * Q_y[IRHO] - u * u_y[IX] - v * u_y[IY] - w * u_y[IZ]);
    nablaT[IZ] = gm1 * gM2 * (invRho * Q_z[IRHOE] - Q[IRHOE] * invRho * invRho * Q_z[IRHO] - u * u_z[IX] - v * u_z[IY] - w * u_z[IZ]);
    double divV = u_x[IX] + u_y[IY] + u_z[IZ];
    F[IRHO][IX] = 0.0;
    F[IRHOU][IX] = mu * (2.0 * u_x[IX] - 2.0 / 3.0 * divV) + beta * divV;
    F[IRHOV][IX] = mu * (u_x[IY] + u_y[IX]);
    F[IRHOW][IX] = mu * (u_x[IZ] + u_z[IX]);
    F[IRHOE][IX] = F[IRHOU][IX] * u + F[IRHOV][IX] * v + F[IRHOW][IX] * w + kappa * nablaT[IX];
    F[IRHO][IY] = 0.0;
    F[IRHOU][IY] = F[IRHOV][IX];
    F[IRHOV][IY] = mu * (2.0 * u_y[IY] - 2.0 / 3.0 * divV) + beta * divV;
    F[IRHOW][IY] = mu * (u_y[IZ] + u_z[IY]);
    F[IRHOE][IY] = F[IRHOU][IY] * u + F[IRHOV][IY] * v + F[IRHOW][IY] * w + kappa * nablaT[IY];
    F[IRHO][IZ] = 0.0;
    F[IRHOU][IZ] = F[IRHOW][IX];
    F[IRHOV][IZ] = F[IRHOW][IY];
    F[IRHOW][IZ] = mu * (2.0 * u_z[IZ] - 2.0 / 3.0 * divV) + beta * divV;
    F[IRHOE][IZ] = F[IRHOU][IZ] * u + F[IRHOV][IZ] * v + F[IRHOW][IZ] * w + kappa * nablaT[IZ];
}

// shared tail of the three ViscousFlux_* routines (Physics_NS.f90:283-303, 338-358, 393-413)
inline void viscousFluxTail(const double* u, const double* u_x, const double* u_y, const double* u_z, const double* nablaT,
                            double mu, double beta, double kappa, double F[NCONS][NDIM]) {
    double divV = u_x[IX] + u_y[IY] + u_z[IZ];
    F[IRHO][IX] = 0.0;
    F[IRHOU][IX] = mu * (2.0 * u_x[IX] - 2.0 / 3.0 * divV) + beta * divV;
    F[IRHOV][IX] = mu * (u_x[IY] + u_y[IX]);
    F[IRHOW][IX] = mu * (u_x[IZ] + u_z[IX]);
    F[IRHOE][IX] = F[IRHOU][IX] * u[IX] + F[IRHOV][IX] * u[IY] + F[IRHOW][IX] * u[IZ] + kappa * nablaT[IX];
    F[IRHO][IY] = 0.0;
    F[IRHOU][IY] = F[IRHOV][IX];
    F[IRHOV][IY] = mu * (2.0 * u_y[IY] - 2.0 / 3.0 * divV) + beta * divV;
    F[IRHOW][IY] = mu * (u_y[IZ] + u_z[IY]);
    F[IRHOE][IY] = F[IRHOU][IY] * u[IX] + F[IRHOV][IY] * u[IY] + F[IRHOW][IY] * u[IZ] + kappa * nablaT[IY];
    F[IRHO][IZ] = 0.0;
    F[IRHOU][IZ] = F[IRHOW][IX];
    F[IRHOV][IZ] = F[IRHOW][IY];
    F[IRHOW][IZ] = mu * (2.0 * u_z[IZ] - 2.0 / 3.0 * divV) + beta * divV;
    F[IRHOE][IZ] = F[IRHOU][IZ] * u[IX] + F[IRHOV][IZ] * u[IY] + F[IRHOW][IZ] * u[IZ] + kappa * nablaT[IZ];
}

// Physics_NS.f90:306-359 (gradients of the entropy variables)
inline void ViscousFlux_ENTROPY(const Oracle& o, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z,
                                double mu, double beta, double kappa, double F[NCONS][NDIM]) {
    double invRho = 1.0 / Q[IRHO];
    double p_div_rho = o.ph.gammaMinus1 * invRho * (Q[IRHOE] - 0.5 * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) * invRho);
    double u[3] = {Q[IRHOU] * invRho, Q[IRHOV] * invRho, Q[IRHOW] * invRho};
    double u_x[3], u_y[3], u_z[3], nablaT[3];
    for (int c = 0; c < 3; ++c) {
        u_x[c] = p_div_rho * (Q_x[IRHOU + c] + u[c] * Q_x[IRHOE]);
        u_y[c] = p_div_rho * (Q_y[IRHOU + c] + u[c] * Q_y[IRHOE]);
        u_z[c] = p_div_rho * (Q_z[IRHOU + c] + u[c] * Q_z[IRHOE]);
    }
    nablaT[IX] = o.ph.gammaM2 * POW2(p_div_rho) * Q_x[IRHOE];
    nablaT[IY] = o.ph.gammaM2 * POW2(p_div_rho) * Q_y[IRHOE];
    nablaT[IZ] = o.ph.gammaM2 * POW2(p_div_rho) * Q_z[IRHOE];
    viscousFluxTail(u, u_x, u_y, u_z, nablaT, mu, beta, kappa, F);
}

// Physics_NS.f90:361-414 (gradients of [rho, u, v, w, T])
inline void ViscousFlux_ENERGY(const Oracle& o, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z,
                               double mu, double beta, double kappa, double F[NCONS][NDIM]) {
    (void)o;
    double invRho = 1.0 / Q[IRHO];
    double u[3] = {Q[IRHOU] * invRho, Q[IRHOV] * invRho, Q[IRHOW] * invRho};
    double u_x[3] = {Q_x[IRHOU], Q_x[IRHOV], Q_x[IRHOW]}, u_y[3] = {Q_y[IRHOU], Q_y[IRHOV], Q_y[IRHOW]}, u_z[3] = {Q_z[IRHOU], Q_z[IRHOV], Q_z[IRHOW]};
    double nablaT[3] = {Q_x[IRHOE], Q_y[IRHOE], Q_z[IRHOE]};
    viscousFluxTail(u, u_x, u_y, u_z, nablaT, mu, beta, kappa, F);
}

// ViscousFlux procedure pointer (SpatialDiscretization.f90:106-148)
inline void ViscousFlux(const Oracle& o, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z,
                        double mu, double beta, double kappa, double F[NCONS][NDIM]) {
    switch (o.ph.gradientVariables) {
        case H3D_GRADVARS_ENTROPY: ViscousFlux_ENTROPY(o, Q, Q_x, Q_y, Q_z, mu, beta, kappa, F); break;
        case H3D_GRADVARS_ENERGY: ViscousFlux_ENERGY(o, Q, Q_x, Q_y, Q_z, mu, beta, kappa, F); break;
        default: ViscousFlux_STATE(o, Q, Q_x, Q_y, Q_z, mu, beta, kappa, F); break;
    }
}

// GetGradients procedure pointer: NSGradientVariables_STATE / _ENTROPY / _ENERGY (VariableConversion_NS.f90:196-262)
inline void GetGradientsAs(const H3dPhysics& ph, int gradVars, const double* Q, double* U);
inline void GetGradients(const Oracle& o, const double* Q, double* U) { GetGradientsAs(o.ph, o.ph.gradientVariables, Q, U); }
inline void GetGradientsAs(const H3dPhysics& ph, int gradVars, const double* Q, double* U) {
    struct { const H3dPhysics& ph; } o{ph};
    switch (gradVars) {
        case H3D_GRADVARS_ENTROPY: {
            double invRho = 1.0 / Q[IRHO];
            double rhoV2 = (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) * invRho;
            double p = o.ph.gammaMinus1 * (Q[IRHOE] - 0.5 * rhoV2);
            double invP = 1.0 / p;
            double invGammaMinus1 = 1.0 / o.ph.gammaMinus1;      // thermodynamics % invGammaMinus1 (FluidData_NS.f90)
            double U0 = (o.ph.gamma - (std::log(p) - o.ph.gamma * std::log(Q[IRHO]))) * invGammaMinus1 - 0.5 * rhoV2 * invP;
            double U4 = -Q[IRHO] * invP;
            U[IRHOU] = Q[IRHOU] * invP; U[IRHOV] = Q[IRHOV] * invP; U[IRHOW] = Q[IRHOW] * invP;
            U[IRHO] = U0; U[IRHOE] = U4;
        } break;
        case H3D_GRADVARS_ENERGY: {
            double invRho = 1.0 / Q[IRHO];
            double rhoV2 = (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) * invRho;
            double p = o.ph.gammaMinus1 * (Q[IRHOE] - 0.5 * rhoV2);
            double U4 = o.ph.gammaM2 * p * invRho;
            double U0 = Q[IRHO];
            U[IRHOU] = Q[IRHOU] * invRho; U[IRHOV] = Q[IRHOV] * invRho; U[IRHOW] = Q[IRHOW] * invRho;
            U[IRHO] = U0; U[IRHOE] = U4;
        } break;
        default: for (int q = 0; q < 5; ++q) U[q] = Q[q]; break;
    }
}

// VariableConversion_NS.f90:50-64, 99-115, 166-186, 147-164
inline double Pressure(const Oracle& o, const double* Q) {
    return o.ph.gammaMinus1 * (Q[4] - 0.5 * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]) / Q[0]);
}
inline double Temperature(const Oracle& o, const double* Q) { return o.ph.gammaM2 * Pressure(o, Q) / Q[0]; }
inline double SutherlandsLaw(const Oracle& o, double T) {
    double tildeT = T * o.ph.T_renorm;
    return (1.0 + o.ph.S_div_Tref) / (tildeT + o.ph.S_div_Tref) * tildeT * std::sqrt(tildeT);
}
inline void get_laminar_mu_kappa(const Oracle& o, const double* Q, double& mu, double& kappa) {
    double T = Temperature(o, Q);
    double suther = SutherlandsLaw(o, T);
    mu = o.ph.mu * suther;
    kappa = mu * o.ph.mu_to_kappa;
}

// VariableConversion_NS.f90:373-392
inline void getVelocityGradients_State(const double* Q, const double* Q_x, const double* Q_y, const double* Q_z, double* U_x, double* U_y, double* U_z) {
    double invRho = 1.0 / Q[IRHO], invRho2 = invRho * invRho;
    double uDivRho[3] = {Q[IRHOU] * invRho2, Q[IRHOV] * invRho2, Q[IRHOW] * invRho2};
    for (int c = 0; c < 3; ++c) {
        U_x[c] = invRho * Q_x[IRHOU + c] - uDivRho[c] * Q_x[IRHO];
        U_y[c] = invRho * Q_y[IRHOU + c] - uDivRho[c] * Q_y[IRHO];
        U_z[c] = invRho * Q_z[IRHOU + c] - uDivRho[c] * Q_z[IRHO];
    }
}

// getVelocityGradients procedure pointer (VariableConversion_NS.f90:373-428, set at :602-617)
inline void getVelocityGradients(const Oracle& o, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z, double* U_x, double* U_y, double* U_z) {
    switch (o.ph.gradientVariables) {
        case H3D_GRADVARS_ENERGY:
            for (int c = 0; c < 3; ++c) { U_x[c] = Q_x[IRHOU + c]; U_y[c] = Q_y[IRHOU + c]; U_z[c] = Q_z[IRHOU + c]; }
            break;
        case H3D_GRADVARS_ENTROPY: {   // as written in the reference (:421-426): U / pDivRho multiplies the energy-variable gradient
            double pDivRho = Pressure(o, Q) / Q[IRHO];
            double U[3] = {Q[IRHOU] / Q[IRHO], Q[IRHOV] / Q[IRHO], Q[IRHOW] / Q[IRHO]};
            for (int c = 0; c < 3; ++c) {
                U_x[c] = pDivRho * Q_x[IRHOU + c] + U[c] / pDivRho * Q_x[IRHOE];
                U_y[c] = pDivRho * Q_y[IRHOU + c] + U[c] / pDivRho * Q_y[IRHOE];
                U_z[c] = pDivRho * Q_z[IRHOU + c] + U[c] / pDivRho * Q_z[IRHOE];
            }
        } break;
        default: getVelocityGradients_State(Q, Q_x, Q_y, Q_z, U_x, U_y, U_z); break;
    }
}

// WALE_ComputeViscosity (LESModels.f90:358-435) and Vreman_ComputeViscosity (:487-546)
inline void getVelocityGradients(const Oracle& o, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z, double* U_x, double* U_y, double* U_z);
inline double WaleViscosity(const Oracle& o, double delta, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z) {
    double U_x[3], U_y[3], U_z[3], gradV[3][3], S[3][3], gradV2[3][3], Sd[3][3];
    getVelocityGradients(o, Q, Q_x, Q_y, Q_z, U_x, U_y, U_z);
    for (int c = 0; c < 3; ++c) { gradV[0][c] = U_x[c]; gradV[1][c] = U_y[c]; gradV[2][c] = U_z[c]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        S[i][j] = 0.5 * (gradV[i][j] + gradV[j][i]);
        gradV2[i][j] = 0;
        for (int k = 0; k < 3; ++k) gradV2[i][j] = gradV2[i][j] + gradV[i][k] * gradV[k][j];
    }
    double divV2 = gradV2[0][0] + gradV2[1][1] + gradV2[2][2];
    double normS = 0.0;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) normS = normS + S[i][j] * S[i][j];      // sum(S*S), column-major
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Sd[i][j] = 0.5 * (gradV2[i][j] + gradV2[j][i]);
    Sd[0][0] = Sd[0][0] - 1.0 / 3.0 * divV2; Sd[1][1] = Sd[1][1] - 1.0 / 3.0 * divV2; Sd[2][2] = Sd[2][2] - 1.0 / 3.0 * divV2;
    double normSd = 0.0;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) normSd = normSd + Sd[i][j] * Sd[i][j];
    double LS = o.ph.smagorinsky_Cs * delta;
    double mu = Q[IRHO] * POW2(LS) * (std::pow(normSd, 3.0 / 2.0) / (std::pow(normS, 5.0 / 2.0) + std::pow(normSd, 5.0 / 4.0)));
    if (normS < 1.0e-8 && normSd < 1.0e-8) mu = 0.0;
    return mu;
}
inline double VremanViscosity(const Oracle& o, double delta, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z) {
    double U_x[3], U_y[3], U_z[3], gradV[3][3], G[3][3];
    getVelocityGradients(o, Q, Q_x, Q_y, Q_z, U_x, U_y, U_z);
    const double delta2 = delta * delta;
    for (int c = 0; c < 3; ++c) { gradV[0][c] = U_x[c]; gradV[1][c] = U_y[c]; gradV[2][c] = U_z[c]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        G[i][j] = 0.0;
        for (int k = 0; k < 3; ++k) G[i][j] = G[i][j] + (gradV[i][k] * gradV[j][k] * delta2);
    }
    double alpha = 0.0;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) alpha = alpha + gradV[i][j] * gradV[i][j];
    double Bbeta = G[0][0] * G[1][1] + G[1][1] * G[2][2] + G[2][2] * G[0][0] - G[0][1] * G[0][1] - G[1][2] * G[1][2] - G[0][2] * G[0][2];
    return alpha > 1.0e-10 ? Q[IRHO] * o.ph.smagorinsky_Cs * std::sqrt(std::fabs(Bbeta) / alpha) : 0.0;
}

// LESModels.f90:256-305 (Smagorinsky_ComputeViscosity) with LESModel_ComputeWallEffect (:189-203); no wall model is the default (:162-165)
inline double SmagorinskyViscosity(const Oracle& o, double delta, double dWall, const double* Q, const double* Q_x, const double* Q_y, const double* Q_z) {
    if (o.ph.les == H3D_LES_WALE) return WaleViscosity(o, delta, Q, Q_x, Q_y, Q_z);        // LESModel % ComputeViscosity is polymorphic
    if (o.ph.les == H3D_LES_VREMAN) return VremanViscosity(o, delta, Q, Q_x, Q_y, Q_z);
    double U_x[3], U_y[3], U_z[3], S[3][3];
    getVelocityGradients(o, Q, Q_x, Q_y, Q_z, U_x, U_y, U_z);
    for (int i = 0; i < 3; ++i) { S[i][0] = U_x[i]; S[i][1] = U_y[i]; S[i][2] = U_z[i]; }
    for (int j = 0; j < 3; ++j) S[0][j] = S[0][j] + U_x[j];
    for (int j = 0; j < 3; ++j) S[1][j] = S[1][j] + U_y[j];
    for (int j = 0; j < 3; ++j) S[2][j] = S[2][j] + U_z[j];
    double sum = 0.0;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) { double sij = 0.5 * S[i][j]; sum = sum + sij * sij; }   // column-major sum(S*S)
    double normS = std::sqrt(2.0 * sum);
    double LS = o.ph.smagorinsky_Cs * delta;
    if (o.ph.les_wall_model == 1) LS = std::fmin(LS, dWall * 0.4);   // K_VONKARMAN = 0.4 (LESModels.f90:31)
    return Q[IRHO] * POW2(LS) * normS;
}

// libs/foundation/Utilities.f90:269-300 (logarithmicMean, Ismail & Roe)
inline double logarithmicMean(double aL, double aR) {
    const double eps = 0.01;
    double xi = aL / aR;
    double f = (xi - 1.0) / (xi + 1.0);
    double u = f * f;
    double FF;
    if (u < eps) FF = 1.0 + (1.0 / 3.0) * u + (1.0 / 5.0) * POW2(u) + (1.0 / 7.0) * (u * u * u);
    else FF = std::log(xi) / (2.0 * f);
    return 0.5 * (aL + aR) / FF;
}

// ---- averaging functions on rotated states (RiemannSolvers_NS.f90:1784-2077)
inline void AveragedStates(const Oracle& o, const double* QL, const double* QR, double pL, double pR, double invRhoL, double invRhoR, double* flux) {
    double uL = invRhoL * QL[IRHOU], uR = invRhoR * QR[IRHOU];
    double vL = invRhoL * QL[IRHOV], vR = invRhoR * QR[IRHOV];
    double wL = invRhoL * QL[IRHOW], wR = invRhoR * QR[IRHOW];
    switch (o.ph.averaging) {
        case H3D_AVG_STANDARD:
            flux[IRHO] = 0.5 * (QL[IRHOU] + QR[IRHOU]);
            flux[IRHOU] = 0.5 * (QL[IRHOU] * uL + QR[IRHOU] * uR + pL + pR);
            flux[IRHOV] = 0.5 * (QL[IRHOU] * vL + QR[IRHOU] * vR);
            flux[IRHOW] = 0.5 * (QL[IRHOU] * wL + QR[IRHOU] * wR);
            flux[IRHOE] = 0.5 * (uL * (QL[IRHOE] + pL) + uR * (QR[IRHOE] + pR));
            break;
        case H3D_AVG_KENNEDYGRUBER: {
            double rho = 0.5 * (QL[IRHO] + QR[IRHO]), u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR), p = 0.5 * (pL + pR);
            double e = 0.5 * (QL[IRHOE] * invRhoL + QR[IRHOE] * invRhoR);
            flux[IRHO] = rho * u; flux[IRHOU] = rho * u * u + p; flux[IRHOV] = rho * u * v; flux[IRHOW] = rho * u * w; flux[IRHOE] = rho * u * e + p * u;
        } break;
        case H3D_AVG_PIROZZOLI: {
            double rho = 0.5 * (QL[IRHO] + QR[IRHO]), u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR), p = 0.5 * (pL + pR);
            double h = 0.5 * ((QL[IRHOE] + pL) * invRhoL + (QR[IRHOE] + pR) * invRhoR);
            flux[IRHO] = rho * u; flux[IRHOU] = rho * u * u + p; flux[IRHOV] = rho * u * v; flux[IRHOW] = rho * u * w; flux[IRHOE] = rho * u * h;
        } break;
        case H3D_AVG_DUCROS:   // :1821-1848
            flux[IRHO] = 0.25 * (QL[IRHO] + QR[IRHO]) * (uL + uR);
            flux[IRHOU] = 0.25 * (QL[IRHOU] + QR[IRHOU]) * (uL + uR) + 0.5 * (pL + pR);
            flux[IRHOV] = 0.25 * (QL[IRHOV] + QR[IRHOV]) * (uL + uR);
            flux[IRHOW] = 0.25 * (QL[IRHOW] + QR[IRHOW]) * (uL + uR);
            flux[IRHOE] = 0.25 * (QL[IRHOE] + pL + QR[IRHOE] + pR) * (uL + uR);
            break;
        case H3D_AVG_MORINISHI: {   // :1850-1886; the enthalpy does not contain the kinetic energy
            const double cp = o.ph.gamma * (1.0 / o.ph.gammaMinus1);   // dimensionless % cp, PhysicsStorage_NS.f90:190
            double hL = cp * pL, hR = cp * pR;
            flux[IRHO] = 0.5 * (QL[IRHOU] + QR[IRHOU]);
            flux[IRHOU] = 0.25 * (QL[IRHOU] + QR[IRHOU]) * (uL + uR) + 0.5 * (pL + pR);
            flux[IRHOV] = 0.25 * (QL[IRHOU] + QR[IRHOU]) * (vL + vR);
            flux[IRHOW] = 0.25 * (QL[IRHOU] + QR[IRHOU]) * (wL + wR);
            flux[IRHOE] = 0.5 * (uL * hL + uR * hR) + 0.25 * (QL[IRHOU] * uL + QR[IRHOU] * uR) * (uL + uR)
                          + 0.25 * (QL[IRHOU] * vL + QR[IRHOU] * vR) * (vL + vR)
                          + 0.25 * (QL[IRHOU] * wL + QR[IRHOU] * wR) * (wL + wR)
                          - 0.25 * (QL[IRHOU] * POW2(uL) + QR[IRHOU] * POW2(uR))
                          - 0.25 * (QL[IRHOU] * POW2(vL) + QR[IRHOU] * POW2(vR))
                          - 0.25 * (QL[IRHOU] * POW2(wL) + QR[IRHOU] * POW2(wR));
        } break;
        case H3D_AVG_ENTROPYCONS: {   // :1964-2026 (Ismail & Roe parameter vector)
            const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
            const double gammaPlus1Div2 = (gamma + 1.0) / 2.0, gammaMinus1Div2 = gm1 / 2.0, gammaDivGammaMinus1 = gamma / gm1, invGamma = 1.0 / gamma;
            double rhoL = QL[IRHO], rhoR = QR[IRHO];
            double zL[5], zR[5], zSum[5];
            zL[4] = std::sqrt(rhoL * pL); zR[4] = std::sqrt(rhoR * pR);
            zL[0] = rhoL / zL[4]; zR[0] = rhoR / zR[4];
            zL[1] = zL[0] * uL; zR[1] = zR[0] * uR;
            zL[2] = zL[0] * vL; zR[2] = zR[0] * vR;
            zL[3] = zL[0] * wL; zR[3] = zR[0] * wR;
            for (int q = 0; q < 5; ++q) zSum[q] = zL[q] + zR[q];
            double invZ1Sum = 1.0 / zSum[0];
            double z1Log = logarithmicMean(zL[0], zR[0]), z5Log = logarithmicMean(zL[4], zR[4]);
            double rho = 0.5 * zSum[0] * z5Log;
            double u = zSum[1] * invZ1Sum, v = zSum[2] * invZ1Sum, w = zSum[3] * invZ1Sum, p = zSum[4] * invZ1Sum;
            double p2 = (gammaPlus1Div2 * z5Log / z1Log + gammaMinus1Div2 * p) * invGamma;
            double h = gammaDivGammaMinus1 * p2 / rho + 0.5 * (POW2(u) + POW2(v) + POW2(w));
            flux[IRHO] = rho * u; flux[IRHOU] = rho * u * u + p; flux[IRHOV] = rho * u * v; flux[IRHOW] = rho * u * w; flux[IRHOE] = rho * u * h;
        } break;
        case H3D_AVG_CHANDRASEKAR: {   // :2028-2077
            double rhoL = QL[IRHO], rhoR = QR[IRHO];
            double betaL = 0.5 * rhoL / pL, betaR = 0.5 * rhoR / pR;
            double betaLog = logarithmicMean(betaL, betaR);
            double rho = logarithmicMean(rhoL, rhoR);
            double u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR);
            double p = 0.5 * (rhoL + rhoR) / (betaL + betaR);
            double h = 0.5 / (betaLog * o.ph.gammaMinus1) - 0.5 * (0.5 * ((POW2(uL) + POW2(vL) + POW2(wL)) + (POW2(uR) + POW2(vR) + POW2(wR))))
                       + p / rho + POW2(u) + POW2(v) + POW2(w);
            flux[IRHO] = rho * u; flux[IRHOU] = rho * u * u + p; flux[IRHOV] = rho * u * v; flux[IRHOW] = rho * u * w; flux[IRHOE] = rho * u * h;
        } break;
        default: for (int q = 0; q < 5; ++q) flux[q] = std::numeric_limits<double>::quiet_NaN();
    }
}

// RiemannSolvers_NS.f90:375-428
inline void CentralRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gm1 = o.ph.gammaMinus1;
    double rhoL = QLeft[0], rhoR = QRight[0], invRhoL = 1.0 / rhoL, invRhoR = 1.0 / rhoR;
    double rhouL = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    double rhovL = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    double rhowL = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    double rhouR = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    double rhovR = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    double rhowR = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    double rhoeL = QLeft[4], rhoeR = QRight[4];
    double rhoV2L = (POW2(rhouL) + POW2(rhovL) + POW2(rhowL)) * invRhoL;
    double rhoV2R = (POW2(rhouR) + POW2(rhovR) + POW2(rhowR)) * invRhoR;
    double pL = gm1 * (rhoeL - 0.5 * rhoV2L), pR = gm1 * (rhoeR - 0.5 * rhoV2R);
    double QLRot[5] = {rhoL, rhouL, rhovL, rhowL, rhoeL}, QRRot[5] = {rhoR, rhouR, rhovR, rhowR, rhoeR};
    AveragedStates(o, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    double f2 = flux[1], f3 = flux[2], f4 = flux[3];
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// RiemannSolvers_NS.f90:1251-1334
inline void LxFRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
    double rhoL = QLeft[0], rhoR = QRight[0], invRhoL = 1.0 / rhoL, invRhoR = 1.0 / rhoR;
    double rhouL = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    double rhouR = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    double rhovL = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    double rhovR = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    double rhowL = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    double rhowR = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    double rhoV2L = (POW2(rhouL) + POW2(rhovL) + POW2(rhowL)) * invRhoL;
    double rhoV2R = (POW2(rhouR) + POW2(rhovR) + POW2(rhowR)) * invRhoR;
    double rhoeL = QLeft[4], rhoeR = QRight[4];
    double pL = gm1 * (rhoeL - 0.5 * rhoV2L), pR = gm1 * (rhoeR - 0.5 * rhoV2R);
    double aL = std::sqrt(gamma * pL * invRhoL), aR = std::sqrt(gamma * pR * invRhoR);
    double lambda = std::fmax(std::fabs(rhouL * invRhoL) + aL, std::fabs(rhouR * invRhoR) + aR);
    double QLRot[5] = {rhoL, rhouL, rhovL, rhowL, rhoeL}, QRRot[5] = {rhoR, rhouR, rhovR, rhowR, rhoeR};
    AveragedStates(o, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    for (int q = 0; q < 5; ++q) { double stab = 0.5 * lambda * (QRRot[q] - QLRot[q]); flux[q] = flux[q] - o.ph.lambdaStab * stab; }
    double f2 = flux[1], f3 = flux[2], f4 = flux[3];
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// RiemannSolvers_NS.f90:1541-1656
inline void RoeRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, double* flux) {
    const double gamma = o.ph.gamma, ds = 1.0;
    double rho = QLeft[0], rhou = QLeft[1], rhov = QLeft[2], rhow = QLeft[3], rhoe = QLeft[4];
    double rhon = QRight[0], rhoun = QRight[1], rhovn = QRight[2], rhown = QRight[3], rhoen = QRight[4];
    double ul = rhou / rho, vl = rhov / rho, wl = rhow / rho;
    double pleft = (gamma - 1.0) * (rhoe - 0.5 / rho * (rhou * rhou + rhov * rhov + rhow * rhow));
    double ur = rhoun / rhon, vr = rhovn / rhon, wr = rhown / rhon;
    double pright = (gamma - 1.0) * (rhoen - 0.5 / rhon * (rhoun * rhoun + rhovn * rhovn + rhown * rhown));
    double ql = nHat[0] * ul + nHat[1] * vl + nHat[2] * wl;
    double qr = nHat[0] * ur + nHat[1] * vr + nHat[2] * wr;
    double hl = 0.5 * (ul * ul + vl * vl + wl * wl) + gamma / (gamma - 1.0) * pleft / rho;
    double hr = 0.5 * (ur * ur + vr * vr + wr * wr) + gamma / (gamma - 1.0) * pright / rhon;
    double rtd = std::sqrt(rho * rhon);
    double betal = rho / (rho + rtd), betar = 1.0 - betal;
    double utd = betal * ul + betar * ur, vtd = betal * vl + betar * vr, wtd = betal * wl + betar * wr, htd = betal * hl + betar * hr;
    double atd2 = (gamma - 1.0) * (htd - 0.5 * (utd * utd + vtd * vtd + wtd * wtd));
    double atd = std::sqrt(atd2);
    double qtd = utd * nHat[0] + vtd * nHat[1] + wtd * nHat[2];
    if (qtd >= 0.0) {
        double dw1 = 0.5 * ((pright - pleft) / atd2 - (qr - ql) * rtd / atd);
        double sp1 = qtd - atd;
        double sp1m = std::fmin(sp1, 0.0);
        double hd1m = ((gamma + 1.0) / 4.0 * atd / rtd) * dw1;
        double eta1 = std::fmax(-std::fabs(sp1) - hd1m, 0.0);
        double udw1 = dw1 * (sp1m - 0.5 * eta1);
        double rql = rho * ql;
        flux[0] = ds * (rql + udw1);
        flux[1] = ds * (rql * ul + pleft * nHat[0] + udw1 * (utd - atd * nHat[0]));
        flux[2] = ds * (rql * vl + pleft * nHat[1] + udw1 * (vtd - atd * nHat[1]));
        flux[3] = ds * (rql * wl + pleft * nHat[2] + udw1 * (wtd - atd * nHat[2]));
        flux[4] = ds * (rql * hl + udw1 * (htd - qtd * atd));
    } else {
        double dw4 = 0.5 * ((pright - pleft) / atd2 + (qr - ql) * rtd / atd);
        double sp4 = qtd + atd;
        double sp4p = std::fmax(sp4, 0.0);
        double hd4 = ((gamma + 1.0) / 4.0 * atd / rtd) * dw4;
        double eta4 = std::fmax(-std::fabs(sp4) + hd4, 0.0);
        double udw4 = dw4 * (sp4p + 0.5 * eta4);
        double rqr = rhon * qr;
        flux[0] = ds * (rqr - udw4);
        flux[1] = ds * (rqr * ur + pright * nHat[0] - udw4 * (utd + atd * nHat[0]));
        flux[2] = ds * (rqr * vr + pright * nHat[1] - udw4 * (vtd + atd * nHat[1]));
        flux[3] = ds * (rqr * wr + pright * nHat[2] - udw4 * (wtd + atd * nHat[2]));
        flux[4] = ds * (rqr * hr - udw4 * (htd + qtd * atd));
    }
}

// RiemannSolvers_NS.f90:1661-1762 (RusanovRiemannSolver): unrotated, smax = max(a + |q|)
inline void RusanovRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, double* flux) {
    const double gamma = o.ph.gamma, ds = 1.0;
    double rho = QLeft[0], rhou = QLeft[1], rhov = QLeft[2], rhow = QLeft[3], rhoe = QLeft[4];
    double rhon = QRight[0], rhoun = QRight[1], rhovn = QRight[2], rhown = QRight[3], rhoen = QRight[4];
    double ul = rhou / rho, vl = rhov / rho, wl = rhow / rho;
    double pleft = (gamma - 1.0) * (rhoe - 0.5 / rho * (rhou * rhou + rhov * rhov + rhow * rhow));
    double ur = rhoun / rhon, vr = rhovn / rhon, wr = rhown / rhon;
    double pright = (gamma - 1.0) * (rhoen - 0.5 / rhon * (rhoun * rhoun + rhovn * rhovn + rhown * rhown));
    double ql = nHat[0] * ul + nHat[1] * vl + nHat[2] * wl;
    double qr = nHat[0] * ur + nHat[1] * vr + nHat[2] * wr;
    double hl = 0.5 * (ul * ul + vl * vl + wl * wl) + gamma / (gamma - 1.0) * pleft / rho;
    double hr = 0.5 * (ur * ur + vr * vr + wr * wr) + gamma / (gamma - 1.0) * pright / rhon;
    double ar2 = (gamma - 1.0) * (hr - 0.5 * (ur * ur + vr * vr + wr * wr));
    double al2 = (gamma - 1.0) * (hl - 0.5 * (ul * ul + vl * vl + wl * wl));
    double ar = std::sqrt(ar2), al = std::sqrt(al2);
    double rql = rho * ql, rqr = rhon * qr;
    flux[0] = ds * (rql + rqr);
    flux[1] = ds * (rql * ul + pleft * nHat[0] + rqr * ur + pright * nHat[0]);
    flux[2] = ds * (rql * vl + pleft * nHat[1] + rqr * vr + pright * nHat[1]);
    flux[3] = ds * (rql * wl + pleft * nHat[2] + rqr * wr + pright * nHat[2]);
    flux[4] = ds * (rql * hl + rqr * hr);
    double smax = std::fmax(ar + std::fabs(qr), al + std::fabs(ql));
    for (int q = 0; q < 5; ++q) flux[q] = (flux[q] - ds * smax * (QRight[q] - QLeft[q])) / 2.0;
}

// RiemannSolvers_NS.f90:430-576 (StdRoeRiemannSolver): rotated, full wave decomposition, Harten / van Leer entropy fix
inline void StdRoeRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
    double QLRot[5], QRRot[5];
    QLRot[0] = QLeft[0]; QRRot[0] = QRight[0];
    QLRot[1] = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    QRRot[1] = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    QLRot[2] = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    QRRot[2] = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    QLRot[3] = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    QRRot[3] = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    QLRot[4] = QLeft[4]; QRRot[4] = QRight[4];
    // getPrimitiveVariables (VariableConversion_NS.f90:266-294): [invRho, u, v, w, p, T, a^2]
    auto prim = [&](const double* U, double* V) {
        double invRho = 1.0 / U[0];
        V[0] = invRho; V[1] = U[1] * invRho; V[2] = U[2] * invRho; V[3] = U[3] * invRho;
        V[4] = gm1 * (U[4] - 0.5 * (V[1] * U[1] + V[2] * U[2] + V[3] * U[3]));
        V[5] = V[4] * o.ph.gammaM2 * invRho;
        V[6] = gamma * V[4] * invRho;
    };
    double VL[7], VR[7];
    prim(QLRot, VL); prim(QRRot, VR);
    double aL = std::sqrt(VL[6]), aR = std::sqrt(VR[6]);
    // getRoeVariables (VariableConversion_NS.f90:323-365)
    double sqrtRhoL = std::sqrt(QLRot[0]), sqrtRhoR = std::sqrt(QRRot[0]);
    double invSumSqrtRhoLR = 1.0 / (sqrtRhoL + sqrtRhoR);
    double HL = (VL[4] + QLRot[4]) * VL[0], HR = (VR[4] + QRRot[4]) * VR[0];
    double u = (sqrtRhoL * VL[1] + sqrtRhoR * VR[1]) * invSumSqrtRhoLR;
    double v = (sqrtRhoL * VL[2] + sqrtRhoR * VR[2]) * invSumSqrtRhoLR;
    double w = (sqrtRhoL * VL[3] + sqrtRhoR * VR[3]) * invSumSqrtRhoLR;
    double H = (sqrtRhoL * HL + sqrtRhoR * HR) * invSumSqrtRhoLR;
    double V2 = POW2(u) + POW2(v) + POW2(w);
    double a = std::sqrt(gm1 * (H - 0.5 * V2));
    double lambda[5] = {u - a, u, u, u, u + a};
    double K[5][5] = {{1.0, u - a, v, w, H - u * a}, {1.0, u, v, w, 0.5 * V2}, {0.0, 0.0, 1.0, 0.0, v}, {0.0, 0.0, 0.0, 1.0, w}, {1.0, u + a, v, w, H + u * a}};   // K[wave][component]
    double dQ[5], alpha[5];
    for (int q = 0; q < 5; ++q) dQ[q] = QRRot[q] - QLRot[q];
    alpha[2] = dQ[2] - v * dQ[0]; alpha[3] = dQ[3] - w * dQ[0];
    dQ[4] = dQ[4] - alpha[2] * v - alpha[3] * w;
    alpha[1] = gm1 * (dQ[0] * (H - u * u) + u * dQ[1] - dQ[4]) / (POW2(a));
    alpha[0] = 0.5 * (dQ[0] * lambda[4] - dQ[1] - a * alpha[1]) / a;
    alpha[4] = dQ[0] - alpha[0] - alpha[1];
    double dLambda = std::fmax((VR[1] - aR) - (VL[1] - aL), 0.0);
    if (std::fabs(lambda[0]) >= 2.0 * dLambda) lambda[0] = std::fabs(lambda[0]);
    else lambda[0] = POW2(lambda[0]) / (4.0 * dLambda) + dLambda;
    dLambda = std::fmax((VR[1] + aR) - (VL[1] + aL), 0.0);
    if (std::fabs(lambda[4]) >= 2.0 * dLambda) lambda[4] = std::fabs(lambda[4]);
    else lambda[4] = POW2(lambda[4]) / (4.0 * dLambda) + dLambda;
    AveragedStates(o, QLRot, QRRot, VL[4], VR[4], VL[0], VR[0], flux);
    if (o.ph.averaging == H3D_AVG_PIROZZOLI || o.ph.averaging == H3D_AVG_KENNEDYGRUBER) lambda[0] = lambda[4];   // Winters et al. correction (:548-558)
    double stab[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; ++i) for (int q = 0; q < 5; ++q) stab[q] = stab[q] + 0.5 * alpha[i] * std::fabs(lambda[i]) * K[i][q];
    for (int q = 0; q < 5; ++q) flux[q] = flux[q] - o.ph.lambdaStab * stab[q];
    double f2 = flux[1], f3 = flux[2], f4 = flux[3];
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// RiemannSolvers_NS.f90:1336-1417 (u_dissRiemannSolver): as Lax-Friedrichs with lambda = max(|uL|, |uR|)
inline void UDissRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gm1 = o.ph.gammaMinus1;
    double rhoL = QLeft[0], rhoR = QRight[0], invRhoL = 1.0 / rhoL, invRhoR = 1.0 / rhoR;
    double rhouL = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    double rhouR = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    double rhovL = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    double rhovR = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    double rhowL = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    double rhowR = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    double rhoV2L = (POW2(rhouL) + POW2(rhovL) + POW2(rhowL)) * invRhoL;
    double rhoV2R = (POW2(rhouR) + POW2(rhovR) + POW2(rhowR)) * invRhoR;
    double rhoeL = QLeft[4], rhoeR = QRight[4];
    double pL = gm1 * (rhoeL - 0.5 * rhoV2L), pR = gm1 * (rhoeR - 0.5 * rhoV2R);
    double lambda = std::fmax(std::fabs(rhouL * invRhoL), std::fabs(rhouR * invRhoR));
    double QLRot[5] = {rhoL, rhouL, rhovL, rhowL, rhoeL}, QRRot[5] = {rhoR, rhouR, rhovR, rhowR, rhoeR};
    AveragedStates(o, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    for (int q = 0; q < 5; ++q) { double stab = 0.5 * lambda * (QRRot[q] - QLRot[q]); flux[q] = flux[q] - o.ph.lambdaStab * stab; }
    double f2 = flux[1], f3 = flux[2], f4 = flux[3];
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// shared tail of the Roe-type solvers (RiemannSolvers_NS.f90:520-572 and the identical blocks of the variants): Harten / van
// Leer entropy fix of waves 1 and 5, averaged flux, Winters correction, stab = sum_i 1/2 alpha_i |lambda_i| K(:,i)
inline void roeTypeTail(const Oracle& o, const double* QLRot, const double* QRRot, double pL, double pR, double invRhoL, double invRhoR,
                        double uL, double uR, double aL, double aR, double u, double v, double w, double H, double a, double V2,
                        const double* alpha, const double* nHat, const double* t1, const double* t2, double* flux) {
    double lambda[5] = {u - a, u, u, u, u + a};
    double K[5][5] = {{1.0, u - a, v, w, H - u * a}, {1.0, u, v, w, 0.5 * V2}, {0.0, 0.0, 1.0, 0.0, v}, {0.0, 0.0, 0.0, 1.0, w}, {1.0, u + a, v, w, H + u * a}};
    double dLambda = std::fmax((uR - aR) - (uL - aL), 0.0);
    if (std::fabs(lambda[0]) >= 2.0 * dLambda) lambda[0] = std::fabs(lambda[0]);
    else lambda[0] = POW2(lambda[0]) / (4.0 * dLambda) + dLambda;
    dLambda = std::fmax((uR + aR) - (uL + aL), 0.0);
    if (std::fabs(lambda[4]) >= 2.0 * dLambda) lambda[4] = std::fabs(lambda[4]);
    else lambda[4] = POW2(lambda[4]) / (4.0 * dLambda) + dLambda;
    AveragedStates(o, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    if (o.ph.averaging == H3D_AVG_PIROZZOLI || o.ph.averaging == H3D_AVG_KENNEDYGRUBER) lambda[0] = lambda[4];
    double stab[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; ++i) for (int q = 0; q < 5; ++q) stab[q] = stab[q] + 0.5 * alpha[i] * std::fabs(lambda[i]) * K[i][q];
    for (int q = 0; q < 5; ++q) flux[q] = flux[q] - o.ph.lambdaStab * stab[q];
    double f2 = flux[1], f3 = flux[2], f4 = flux[3];
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

// RiemannSolvers_NS.f90:721-861 (RoePikeRiemannSolver): as the standard Roe solver with the wave strengths from primitive jumps
inline void RoePikeRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
    double QLRot[5], QRRot[5];
    QLRot[0] = QLeft[0]; QRRot[0] = QRight[0];
    QLRot[1] = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    QRRot[1] = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    QLRot[2] = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    QRRot[2] = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    QLRot[3] = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    QRRot[3] = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    QLRot[4] = QLeft[4]; QRRot[4] = QRight[4];
    auto prim = [&](const double* U, double* V) {
        double invRho = 1.0 / U[0];
        V[0] = invRho; V[1] = U[1] * invRho; V[2] = U[2] * invRho; V[3] = U[3] * invRho;
        V[4] = gm1 * (U[4] - 0.5 * (V[1] * U[1] + V[2] * U[2] + V[3] * U[3]));
        V[6] = gamma * V[4] * invRho;
    };
    double VL[7], VR[7];
    prim(QLRot, VL); prim(QRRot, VR);
    double aL = std::sqrt(VL[6]), aR = std::sqrt(VR[6]);
    double sqrtRhoL = std::sqrt(QLRot[0]), sqrtRhoR = std::sqrt(QRRot[0]);
    double invSumSqrtRhoLR = 1.0 / (sqrtRhoL + sqrtRhoR);
    double HL = (VL[4] + QLRot[4]) * VL[0], HR = (VR[4] + QRRot[4]) * VR[0];
    double rho = sqrtRhoL * sqrtRhoR;
    double u = (sqrtRhoL * VL[1] + sqrtRhoR * VR[1]) * invSumSqrtRhoLR;
    double v = (sqrtRhoL * VL[2] + sqrtRhoR * VR[2]) * invSumSqrtRhoLR;
    double w = (sqrtRhoL * VL[3] + sqrtRhoR * VR[3]) * invSumSqrtRhoLR;
    double H = (sqrtRhoL * HL + sqrtRhoR * HR) * invSumSqrtRhoLR;
    double V2 = POW2(u) + POW2(v) + POW2(w);
    double a = std::sqrt(gm1 * (H - 0.5 * V2));
    double alpha[5];
    alpha[0] = ((VR[4] - VL[4]) - rho * a * (VR[1] - VL[1])) / (2.0 * a * a);
    alpha[1] = (QRight[0] - QLeft[0]) - (VR[4] - VL[4]) / (a * a);
    alpha[2] = rho * (VR[2] - VL[2]);
    alpha[3] = rho * (VR[3] - VL[3]);
    alpha[4] = ((VR[4] - VL[4]) + rho * a * (VR[1] - VL[1])) / (2.0 * a * a);
    roeTypeTail(o, QLRot, QRRot, VL[4], VR[4], VL[0], VR[0], VL[1], VR[1], aL, aR, u, v, w, H, a, V2, alpha, nHat, t1, t2, flux);
}

// RiemannSolvers_NS.f90:863-1058 (LowDissipationRoeRiemannSolver): velocity jumps scaled by z = min(1, max(M_L, M_R))
inline void LowDissipationRoeRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
    double rhoL = QLeft[0], rhoR = QRight[0], invRhoL = 1.0 / rhoL, invRhoR = 1.0 / rhoR;
    double sqrtRhoL = std::sqrt(rhoL), sqrtRhoR = std::sqrt(rhoR);
    double invSqrtRhoL = 1.0 / sqrtRhoL, invSqrtRhoR = 1.0 / sqrtRhoR;
    double invSumSqrtRhoLR = 1.0 / (sqrtRhoL + sqrtRhoR);
    double rhouL = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    double rhouR = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    double rhovL = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    double rhovR = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    double rhowL = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    double rhowR = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    double rhoeL = QLeft[4], rhoeR = QRight[4];
    double uL = rhouL * invRhoL, uR = rhouR * invRhoR, vL = rhovL * invRhoL, vR = rhovR * invRhoR, wL = rhowL * invRhoL, wR = rhowR * invRhoR;
    double rhoV2L = (POW2(uL) + POW2(vL) + POW2(wL)) * rhoL, rhoV2R = (POW2(uR) + POW2(vR) + POW2(wR)) * rhoR;
    double rhoHL = gamma * rhoeL - 0.5 * gm1 * rhoV2L, rhoHR = gamma * rhoeR - 0.5 * gm1 * rhoV2R;
    double pL = gm1 * (rhoeL - 0.5 * rhoV2L), pR = gm1 * (rhoeR - 0.5 * rhoV2R);
    double aL = std::sqrt(gamma * pL * invRhoL), aR = std::sqrt(gamma * pR * invRhoR);
    double rho = sqrtRhoL * sqrtRhoR;
    double u = (invSqrtRhoL * rhouL + invSqrtRhoR * rhouR) * invSumSqrtRhoLR;
    double v = (invSqrtRhoL * rhovL + invSqrtRhoR * rhovR) * invSumSqrtRhoLR;
    double w = (invSqrtRhoL * rhowL + invSqrtRhoR * rhowR) * invSumSqrtRhoLR;
    double H = (invSqrtRhoL * rhoHL + invSqrtRhoR * rhoHR) * invSumSqrtRhoLR;
    double V2abs = POW2(u) + POW2(v) + POW2(w);
    double a = std::sqrt(gm1 * (H - 0.5 * V2abs));
    double ML = std::fabs(uL) / aL, MR = std::fabs(uR) / aR;
    double z = std::fmin(1.0, std::fmax(ML, MR));
    double du = z * (uR - uL), dv = z * (vR - vL), dw = z * (wR - wL), dp = pR - pL;
    double alpha[5];
    alpha[0] = (dp - rho * a * du) / (2.0 * a * a);
    alpha[1] = (rhoR - rhoL) - dp / (a * a);
    alpha[2] = rho * dv;
    alpha[3] = rho * dw;
    alpha[4] = (dp + rho * a * du) / (2.0 * a * a);
    double QLRot[5] = {rhoL, rhouL, rhovL, rhowL, rhoeL}, QRRot[5] = {rhoR, rhouR, rhovR, rhowR, rhoeR};
    roeTypeTail(o, QLRot, QRRot, pL, pR, invRhoL, invRhoR, uL, uR, aL, aR, u, v, w, H, a, V2abs, alpha, nHat, t1, t2, flux);
}

// RiemannSolvers_NS.f90:578-719 (MatrixDissipationRiemannSolver): entropy-variable jump, Chandrasekar mean state
inline void MatrixDissipationRiemannSolver(const Oracle& o, const double* QLeft, const double* QRight, const double* nHat, const double* t1, const double* t2, double* flux) {
    const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
    const double invGamma = 1.0 / gamma, cp = gamma / gm1, gammaMinus1Div2g = gm1 / (2.0 * gamma), invGammaMinus1 = 1.0 / gm1;
    double QLRot[5], QRRot[5];
    QLRot[0] = QLeft[0]; QRRot[0] = QRight[0];
    QLRot[1] = QLeft[1] * nHat[0] + QLeft[2] * nHat[1] + QLeft[3] * nHat[2];
    QRRot[1] = QRight[1] * nHat[0] + QRight[2] * nHat[1] + QRight[3] * nHat[2];
    QLRot[2] = QLeft[1] * t1[0] + QLeft[2] * t1[1] + QLeft[3] * t1[2];
    QRRot[2] = QRight[1] * t1[0] + QRight[2] * t1[1] + QRight[3] * t1[2];
    QLRot[3] = QLeft[1] * t2[0] + QLeft[2] * t2[1] + QLeft[3] * t2[2];
    QRRot[3] = QRight[1] * t2[0] + QRight[2] * t2[1] + QRight[3] * t2[2];
    QLRot[4] = QLeft[4]; QRRot[4] = QRight[4];
    auto entropyVars = [&](const double* Q, double* U) {   // NSGradientVariables_ENTROPY, VariableConversion_NS.f90:211-237
        double invRho = 1.0 / Q[0];
        double rhoV2 = (POW2(Q[1]) + POW2(Q[2]) + POW2(Q[3])) * invRho;
        double p = gm1 * (Q[4] - 0.5 * rhoV2);
        double invP = 1.0 / p;
        U[0] = (gamma - (std::log(p) - gamma * std::log(Q[0]))) * invGammaMinus1 - 0.5 * rhoV2 * invP;
        U[1] = Q[1] * invP; U[2] = Q[2] * invP; U[3] = Q[3] * invP; U[4] = -Q[0] * invP;
    };
    double EVL[5], EVR[5];
    entropyVars(QLRot, EVL); entropyVars(QRRot, EVR);
    double invRhoL = 1.0 / QLRot[0], invRhoR = 1.0 / QRRot[0];
    double uL = QLRot[1] * invRhoL, uR = QRRot[1] * invRhoR, vL = QLRot[2] * invRhoL, vR = QRRot[2] * invRhoR, wL = QLRot[3] * invRhoL, wR = QRRot[3] * invRhoR;
    double vtotL = uL * uL + vL * vL + wL * wL, vtotR = uR * uR + vR * vR + wR * wR;
    double pL = gm1 * (QLRot[4] - 0.5 * QLRot[0] * vtotL), pR = gm1 * (QRRot[4] - 0.5 * QRRot[0] * vtotR);
    double betaL = -0.5 * EVL[4], betaR = -0.5 * EVR[4];
    double betaLogMean = logarithmicMean(betaL, betaR), rhoLogMean = logarithmicMean(QLRot[0], QRRot[0]);
    double pMean = 0.5 * (QLRot[0] + QRRot[0]) / (betaL + betaR);
    double a_bar = std::sqrt(gamma * pMean / rhoLogMean);
    double uMean = 0.5 * (uL + uR), vMean = 0.5 * (vL + vR), wMean = 0.5 * (wL + wR);
    double V2abs = 2.0 * (POW2(uMean) + POW2(vMean) + POW2(wMean)) - 0.5 * (vtotL + vtotR);
    double h_bar = 0.5 * (cp / betaLogMean + V2abs);
    double lambda[5] = {std::fabs(uMean - a_bar), std::fabs(uMean), std::fabs(uMean), std::fabs(uMean), std::fabs(uMean + a_bar)};
    // R1(component, wave)
    double R1[5][5] = {{1.0, 1.0, 0.0, 0.0, 1.0},
                       {uMean - a_bar, uMean, 0.0, 0.0, uMean + a_bar},
                       {vMean, vMean, 1.0, 0.0, vMean},
                       {wMean, wMean, 0.0, 1.0, wMean},
                       {h_bar - uMean * a_bar, 0.5 * V2abs, vMean, wMean, h_bar + uMean * a_bar}};
    double T[5];
    T[0] = 0.5 * rhoLogMean * invGamma; T[1] = 2.0 * gammaMinus1Div2g * rhoLogMean; T[2] = pMean; T[3] = pMean; T[4] = T[0];
    AveragedStates(o, QLRot, QRRot, pL, pR, invRhoL, invRhoR, flux);
    if (o.ph.averaging == H3D_AVG_PIROZZOLI || o.ph.averaging == H3D_AVG_KENNEDYGRUBER) lambda[0] = lambda[4];
    double stab[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) for (int k = 0; k < 5; ++k)
        stab[i] = stab[i] + 0.5 * R1[i][j] * lambda[j] * T[j] * R1[k][j] * (EVR[k] - EVL[k]);
    for (int q = 0; q < 5; ++q) flux[q] = flux[q] - o.ph.lambdaStab * stab[q];
    double f2 = flux[1], f3 = flux[2], f4 = flux[3];
    for (int c = 0; c < 3; ++c) flux[1 + c] = nHat[c] * f2 + t1[c] * f3 + t2[c] * f4;
}

inline void RiemannSolver(const Oracle& o, const double* QL, const double* QR, const double* nHat, const double* t1, const double* t2, double* flux) {
    switch (o.ph.riemann) {
        case H3D_RIEMANN_ROEPIKE: RoePikeRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        case H3D_RIEMANN_LOWDISSROE: LowDissipationRoeRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        case H3D_RIEMANN_MATRIXDISS: MatrixDissipationRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        case H3D_RIEMANN_RUSANOV: RusanovRiemannSolver(o, QL, QR, nHat, flux); break;
        case H3D_RIEMANN_STDROE: StdRoeRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        case H3D_RIEMANN_UDISS: UDissRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        case H3D_RIEMANN_ROE: RoeRiemannSolver(o, QL, QR, nHat, flux); break;
        case H3D_RIEMANN_LXF: LxFRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        case H3D_RIEMANN_CENTRAL: CentralRiemannSolver(o, QL, QR, nHat, t1, t2, flux); break;
        default: for (int q = 0; q < 5; ++q) flux[q] = std::numeric_limits<double>::quiet_NaN();
    }
}

// ---- two-point fluxes (RiemannSolvers_NS.f90:2085-2145, 2296-2367, 2369-2437)
inline void TwoPointFlux(const Oracle& o, const double* QL, const double* QR, const double* JaL, const double* JaR, double* fSharp) {
    const double gm1 = o.ph.gammaMinus1;
    double invRhoL = 1.0 / QL[IRHO], invRhoR = 1.0 / QR[IRHO];
    double uL = invRhoL * QL[IRHOU], uR = invRhoR * QR[IRHOU];
    double vL = invRhoL * QL[IRHOV], vR = invRhoR * QR[IRHOV];
    double wL = invRhoL * QL[IRHOW], wR = invRhoR * QR[IRHOW];
    double pL = gm1 * (QL[IRHOE] - 0.5 * (QL[IRHOU] * uL + QL[IRHOV] * vL + QL[IRHOW] * wL));
    double pR = gm1 * (QR[IRHOE] - 0.5 * (QR[IRHOU] * uR + QR[IRHOV] * vR + QR[IRHOW] * wR));
    double f[5], g[5], h[5], Ja[3];
    if (o.ph.averaging == H3D_AVG_STANDARD) {
        for (int c = 0; c < 3; ++c) Ja[c] = (JaL[c] + JaR[c]);
        f[IRHO] = (QL[IRHOU] + QR[IRHOU]);
        f[IRHOU] = (QL[IRHOU] * uL + QR[IRHOU] * uR + pL + pR);
        f[IRHOV] = (QL[IRHOU] * vL + QR[IRHOU] * vR);
        f[IRHOW] = (QL[IRHOU] * wL + QR[IRHOU] * wR);
        f[IRHOE] = (uL * (QL[IRHOE] + pL) + uR * (QR[IRHOE] + pR));
        g[IRHO] = (QL[IRHOV] + QR[IRHOV]);
        g[IRHOU] = (QL[IRHOV] * uL + QR[IRHOV] * uR);
        g[IRHOV] = (QL[IRHOV] * vL + QR[IRHOV] * vR + pL + pR);
        g[IRHOW] = (QL[IRHOV] * wL + QR[IRHOV] * wR);
        g[IRHOE] = (vL * (QL[IRHOE] + pL) + vR * (QR[IRHOE] + pR));
        h[IRHO] = (QL[IRHOW] + QR[IRHOW]);
        h[IRHOU] = (QL[IRHOW] * uL + QR[IRHOW] * uR);
        h[IRHOV] = (QL[IRHOW] * vL + QR[IRHOW] * vR);
        h[IRHOW] = (QL[IRHOW] * wL + QR[IRHOW] * wR + pL + pR);
        h[IRHOE] = (wL * (QL[IRHOE] + pL) + wR * (QR[IRHOE] + pR));
        for (int q = 0; q < 5; ++q) fSharp[q] = 0.25 * (f[q] * Ja[IX] + g[q] * Ja[IY] + h[q] * Ja[IZ]);
        return;
    }
    double rho = 0.5 * (QL[IRHO] + QR[IRHO]), u = 0.5 * (uL + uR), v = 0.5 * (vL + vR), w = 0.5 * (wL + wR), p = 0.5 * (pL + pR);
    for (int c = 0; c < 3; ++c) Ja[c] = 0.5 * (JaL[c] + JaR[c]);
    if (o.ph.averaging == H3D_AVG_KENNEDYGRUBER) {
        double e = 0.5 * (QL[IRHOE] * invRhoL + QR[IRHOE] * invRhoR);
        f[IRHO] = rho * u; f[IRHOU] = rho * u * u + p; f[IRHOV] = rho * u * v; f[IRHOW] = rho * u * w; f[IRHOE] = rho * u * e + p * u;
        g[IRHO] = rho * v; g[IRHOU] = rho * v * u; g[IRHOV] = rho * v * v + p; g[IRHOW] = rho * v * w; g[IRHOE] = rho * v * e + p * v;
        h[IRHO] = rho * w; h[IRHOU] = rho * w * u; h[IRHOV] = rho * w * v; h[IRHOW] = rho * w * w + p; h[IRHOE] = rho * w * e + p * w;
    } else if (o.ph.averaging == H3D_AVG_PIROZZOLI) {
        double hh = 0.5 * ((QL[IRHOE] + pL) * invRhoL + (QR[IRHOE] + pR) * invRhoR);
        f[IRHO] = rho * u; f[IRHOU] = rho * u * u + p; f[IRHOV] = rho * u * v; f[IRHOW] = rho * u * w; f[IRHOE] = rho * u * hh;
        g[IRHO] = rho * v; g[IRHOU] = rho * v * u; g[IRHOV] = rho * v * v + p; g[IRHOW] = rho * v * w; g[IRHOE] = rho * v * hh;
        h[IRHO] = rho * w; h[IRHOU] = rho * w * u; h[IRHOV] = rho * w * v; h[IRHOW] = rho * w * w + p; h[IRHOE] = rho * w * hh;
    } else if (o.ph.averaging == H3D_AVG_DUCROS) {   // :2231-2294
        const double QLm[3] = {QL[IRHOU], QL[IRHOV], QL[IRHOW]}, QRm[3] = {QR[IRHOU], QR[IRHOV], QR[IRHOW]};
        const double velSum[3] = {uL + uR, vL + vR, wL + wR};
        double* F[3] = {f, g, h};
        for (int d = 0; d < 3; ++d) {
            F[d][IRHO] = 0.25 * (QL[IRHO] + QR[IRHO]) * velSum[d];
            for (int c = 0; c < 3; ++c) F[d][IRHOU + c] = 0.25 * (QLm[c] + QRm[c]) * velSum[d];
            F[d][IRHOU + d] = 0.25 * (QLm[d] + QRm[d]) * velSum[d] + 0.5 * (pL + pR);
            F[d][IRHOE] = 0.25 * (QL[IRHOE] + pL + QR[IRHOE] + pR) * velSum[d];
        }
    } else if (o.ph.averaging == H3D_AVG_MORINISHI) {   // :2147-2229
        const double cp = o.ph.gamma * (1.0 / o.ph.gammaMinus1);
        const double hL = cp * pL, hR = cp * pR;
        const double QLm[3] = {QL[IRHOU], QL[IRHOV], QL[IRHOW]}, QRm[3] = {QR[IRHOU], QR[IRHOV], QR[IRHOW]};
        const double velL[3] = {uL, vL, wL}, velR[3] = {uR, vR, wR};
        double* F[3] = {f, g, h};
        for (int d = 0; d < 3; ++d) {
            F[d][IRHO] = 0.5 * (QLm[d] + QRm[d]);
            for (int c = 0; c < 3; ++c) F[d][IRHOU + c] = 0.25 * (QLm[d] + QRm[d]) * (velL[c] + velR[c]);
            F[d][IRHOU + d] = 0.25 * (QLm[d] + QRm[d]) * (velL[d] + velR[d]) + 0.5 * (pL + pR);
            F[d][IRHOE] = 0.5 * (velL[d] * hL + velR[d] * hR) + 0.25 * (QLm[d] * uL + QRm[d] * uR) * (uL + uR)
                          + 0.25 * (QLm[d] * vL + QRm[d] * vR) * (vL + vR)
                          + 0.25 * (QLm[d] * wL + QRm[d] * wR) * (wL + wR)
                          - 0.25 * (QLm[d] * POW2(uL) + QRm[d] * POW2(uR))
                          - 0.25 * (QLm[d] * POW2(vL) + QRm[d] * POW2(vR))
                          - 0.25 * (QLm[d] * POW2(wL) + QRm[d] * POW2(wR));
        }
    } else if (o.ph.averaging == H3D_AVG_ENTROPYCONS || o.ph.averaging == H3D_AVG_CHANDRASEKAR) {   // :2439-2558, :2560-2641
        double rhoL = QL[IRHO], rhoR = QR[IRHO], rhoM, uM, vM, wM, pM, hM;
        if (o.ph.averaging == H3D_AVG_ENTROPYCONS) {
            const double gamma = o.ph.gamma;
            const double gammaPlus1Div2 = (gamma + 1.0) / 2.0, gammaMinus1Div2 = gm1 / 2.0, gammaDivGammaMinus1 = gamma / gm1, invGamma = 1.0 / gamma;
            double zL[5], zR[5], zAv[5];
            zL[4] = std::sqrt(rhoL * pL); zR[4] = std::sqrt(rhoR * pR);
            zL[0] = rhoL / zL[4]; zR[0] = rhoR / zR[4];
            zL[1] = zL[0] * uL; zR[1] = zR[0] * uR;
            zL[2] = zL[0] * vL; zR[2] = zR[0] * vR;
            zL[3] = zL[0] * wL; zR[3] = zR[0] * wR;
            for (int q = 0; q < 5; ++q) zAv[q] = 0.5 * (zL[q] + zR[q]);
            double invZ1Av = 1.0 / zAv[0];
            double z1Log = logarithmicMean(zL[0], zR[0]), z5Log = logarithmicMean(zL[4], zR[4]);
            rhoM = zAv[0] * z5Log;
            uM = zAv[1] * invZ1Av; vM = zAv[2] * invZ1Av; wM = zAv[3] * invZ1Av; pM = zAv[4] * invZ1Av;
            double p2 = (gammaPlus1Div2 * z5Log / z1Log + gammaMinus1Div2 * pM) * invGamma;
            hM = gammaDivGammaMinus1 * p2 / rhoM + 0.5 * (POW2(uM) + POW2(vM) + POW2(wM));
        } else {
            double betaL = 0.5 * rhoL / pL, betaR = 0.5 * rhoR / pR;
            double betaLog = logarithmicMean(betaL, betaR);
            rhoM = logarithmicMean(rhoL, rhoR);
            uM = 0.5 * (uL + uR); vM = 0.5 * (vL + vR); wM = 0.5 * (wL + wR);
            pM = 0.5 * (rhoL + rhoR) / (betaL + betaR);
            hM = 0.5 / (betaLog * gm1) - 0.5 * (0.5 * ((POW2(uL) + POW2(vL) + POW2(wL)) + (POW2(uR) + POW2(vR) + POW2(wR))))
                 + pM / rhoM + POW2(uM) + POW2(vM) + POW2(wM);
        }
        f[IRHO] = rhoM * uM; f[IRHOU] = rhoM * uM * uM + pM; f[IRHOV] = rhoM * uM * vM; f[IRHOW] = rhoM * uM * wM; f[IRHOE] = rhoM * uM * hM;
        g[IRHO] = rhoM * vM; g[IRHOU] = rhoM * vM * uM; g[IRHOV] = rhoM * vM * vM + pM; g[IRHOW] = rhoM * vM * wM; g[IRHOE] = rhoM * vM * hM;
        h[IRHO] = rhoM * wM; h[IRHOU] = rhoM * wM * uM; h[IRHOV] = rhoM * wM * vM; h[IRHOW] = rhoM * wM * wM + pM; h[IRHOE] = rhoM * wM * hM;
    } else {
        for (int q = 0; q < 5; ++q) f[q] = g[q] = h[q] = std::numeric_limits<double>::quiet_NaN();
    }
    for (int q = 0; q < 5; ++q) fSharp[q] = f[q] * Ja[IX] + g[q] * Ja[IY] + h[q] * Ja[IZ];
}

// ------------------------------------------------------------------------------------------------
//  Boundary conditions (libs/physics/common): FlowState, FlowGradVars, FlowNeumann.
//  Zone parameters P[16] (prepared by the host from the control file, see include/h3d_gpu.h):
//    no-slip / free-slip wall: P[0..2] vWall, P[3] wallType (0 adiabatic, 1 isothermal), P[4] Twall,
//                              P[5] refValues%T*gammaM2*gammaMinus1 (no-slip) or refValues%T*gammaM2 (free-slip), P[6] eWall
//    inflow: P[0] rho, P[1..3] u,v,w, P[4] p        outflow: P[4] pExt
// ------------------------------------------------------------------------------------------------
inline void BC_FlowState(const Oracle& o, int zone, const double* nHat, double* Q) {
    const double* P = &o.bcParams[16 * zone];
    const double gamma = o.ph.gamma, gm1 = o.ph.gammaMinus1;
    switch (o.bcType[zone]) {
        case H3D_BC_NOSLIPWALL: {   // NoSlipWallBC.f90:265-296
            Q[IRHOU] = 2.0 * Q[IRHO] * P[0] - Q[IRHOU]; Q[IRHOV] = 2.0 * Q[IRHO] * P[1] - Q[IRHOV]; Q[IRHOW] = 2.0 * Q[IRHO] * P[2] - Q[IRHOW];
            Q[IRHOE] = Q[IRHOE] + P[3] * (Q[IRHO] * P[4] / P[5] - Q[IRHOE]);
        } break;
        case H3D_BC_FREESLIPWALL: { // FreeSlipWallBC.f90:249-283
            double qNorm = nHat[IX] * Q[IRHOU] + nHat[IY] * Q[IRHOV] + nHat[IZ] * Q[IRHOW];
            Q[IRHOU] = Q[IRHOU] - 2.0 * qNorm * nHat[0]; Q[IRHOV] = Q[IRHOV] - 2.0 * qNorm * nHat[1]; Q[IRHOW] = Q[IRHOW] - 2.0 * qNorm * nHat[2];
            double pressure_aux = Q[IRHO] * P[4] / P[5];
            Q[IRHOE] = Q[IRHOE] + P[3] * (pressure_aux / gm1 + 0.5 * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) / Q[IRHO] - Q[IRHOE]);
        } break;
        case H3D_BC_INFLOW: {       // InflowBC.f90:363-406 (TurbIntensity = 0)
            double u = P[1], v = P[2], w = P[3];
            Q[0] = P[0]; Q[1] = Q[0] * u; Q[2] = Q[0] * v; Q[3] = Q[0] * w;
            Q[4] = P[4] / (gamma - 1.0) + 0.5 * Q[0] * (u * u + v * v + w * w);
        } break;
        case H3D_BC_OUTFLOW: {      // OutflowBC.f90:226-288
            const double pExt = P[4];
            double qDotN = (nHat[0] * Q[1] + nHat[1] * Q[2] + nHat[2] * Q[3]) / Q[0];
            double qTanx = Q[1] / Q[0] - qDotN * nHat[0], qTany = Q[2] / Q[0] - qDotN * nHat[1], qTanz = Q[3] / Q[0] - qDotN * nHat[2];
            double p = gm1 * (Q[4] - 0.5 * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]) / Q[0]);
            double a2 = gamma * p / Q[0];
            double a = std::sqrt(a2);
            double normalMachNo = std::fabs(qDotN / a);
            if (normalMachNo <= 1.0) {
                double rPlus = qDotN + 2.0 * a / gm1;
                double entropyConstant = p - a2 * Q[0];
                double rho = -(entropyConstant - pExt) / a2;
                a = std::sqrt(gamma * pExt / rho);
                qDotN = rPlus - 2.0 * a / gm1;
                double u = qTanx + qDotN * nHat[0], v = qTany + qDotN * nHat[1], w = qTanz + qDotN * nHat[2];
                Q[0] = rho; Q[1] = rho * u; Q[2] = rho * v; Q[3] = rho * w;
                Q[4] = pExt / gm1 + 0.5 * rho * (u * u + v * v + w * w);
            }
        } break;
        default: break;
    }
}

// u_star for BR1_ComputeBoundaryFlux (EllipticBR1.f90:686-736): GradVarsForEqn -> FlowGradVars, STATE gradient variables
inline void BC_FlowGradVars(const Oracle& o, int zone, const double* nHat, const double* Q, double* U) {
    const double* P = &o.bcParams[16 * zone];
    switch (o.bcType[zone]) {
        case H3D_BC_NOSLIPWALL: {   // NoSlipWallBC.f90:298-337 ; U(IRHO) keeps the incoming (interior) value
            double invRho = 1.0 / Q[IRHO];
            double e_int = invRho * (Q[IRHOE] - 0.5 * invRho * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])));
            double U1 = U[IRHO], Q_aux[5];
            Q_aux[IRHO] = Q[IRHO];
            Q_aux[IRHOU] = Q[IRHO] * P[0]; Q_aux[IRHOV] = Q[IRHO] * P[1]; Q_aux[IRHOW] = Q[IRHO] * P[2];
            Q_aux[IRHOE] = Q[IRHO] * ((1.0 - P[3]) * e_int + P[3] * P[6] + 0.5 * (P[0] * P[0] + P[1] * P[1] + P[2] * P[2]));
            GetGradients(o, Q_aux, U);
            U[IRHO] = U1;
        } break;
        case H3D_BC_FREESLIPWALL: { // FreeSlipWallBC.f90:285-312
            double Q_aux[5];
            Q_aux[IRHO] = Q[IRHO]; Q_aux[IRHOU] = Q[IRHOU]; Q_aux[IRHOV] = Q[IRHOV]; Q_aux[IRHOW] = Q[IRHOW];
            Q_aux[IRHOE] = Q[IRHOE] + P[3] * (Q[IRHO] * P[6] + 0.5 * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) / Q[IRHO] - Q[IRHOE]);
            GetGradients(o, Q_aux, U);
        } break;
        default: {                  // GenericBC_FlowGradVars (GenericBoundaryConditionClass.f90:243-278): gradient variables of the external state
            double Q_aux[5], U_aux[5];
            for (int q = 0; q < 5; ++q) Q_aux[q] = Q[q];
            BC_FlowState(o, zone, nHat, Q_aux);
            GetGradients(o, Q_aux, U_aux);
            for (int q = 0; q < 5; ++q) U[q] = 0.5 * (U_aux[q] + U[q]);
        } break;
    }
}

// FlowNeumann fix-up of the boundary viscous flux
inline void BC_FlowNeumann(const Oracle& o, int zone, const double* Q, double* flux) {
    const double* P = &o.bcParams[16 * zone];
    switch (o.bcType[zone]) {
        case H3D_BC_NOSLIPWALL: {   // NoSlipWallBC.f90:339-372
            double invRho = 1.0 / Q[IRHO], u = invRho * Q[IRHOU], v = invRho * Q[IRHOV], w = invRho * Q[IRHOW];
            double viscWork = u * flux[IRHOU] + v * flux[IRHOV] + w * flux[IRHOW];
            double heatFlux = flux[IRHOE] - viscWork;
            flux[IRHO] = 0.0;
            flux[IRHOE] = (P[0] * flux[IRHOU] + P[1] * flux[IRHOV] + P[2] * flux[IRHOW]) + P[3] * heatFlux;
        } break;
        case H3D_BC_FREESLIPWALL: { // FreeSlipWallBC.f90:314-351
            double viscWork = (flux[IRHOU] * Q[IRHOU] + flux[IRHOV] * Q[IRHOV] + flux[IRHOW] * Q[IRHOW]) / Q[IRHO];
            double heatFlux = flux[IRHOE] - viscWork;
            flux[IRHO] = 0.0; flux[IRHOU] = 0.0; flux[IRHOV] = 0.0; flux[IRHOW] = 0.0;
            flux[IRHOE] = P[3] * heatFlux;
        } break;
        case H3D_BC_INFLOW: case H3D_BC_OUTFLOW: for (int q = 0; q < 5; ++q) flux[q] = 0.0; break;   // InflowBC.f90:408-427, OutflowBC.f90:290-305
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------
//  Index helpers
// ------------------------------------------------------------------------------------------------
struct Idx {
    int n;
    inline size_t node(int e, int i, int j, int k) const { return ((size_t)e * n * n * n + (size_t)(k * n + j) * n + i); }
    inline size_t fnode(int f, int side, int i, int j) const { return (((size_t)f * 2 + side) * n * n + (size_t)j * n + i); }
    inline size_t gnode(int f, int i, int j) const { return ((size_t)f * n * n + (size_t)j * n + i); }
};

// element-trace index (a,b) on local face lf <-> volume indices
static const int axisMap[6][2] = {{0, 2}, {0, 2}, {0, 1}, {1, 2}, {0, 1}, {1, 2}};

// ------------------------------------------------------------------------------------------------
//  HexMesh_ProlongSolutionToFaces (libs/mesh/HexMesh.f90:948-1036) -> HexElement_ProlongSolutionToFaces
//  (HexElementClass.f90:233-302) -> Face_AdaptSolutionToFace (FaceClass.f90:281-381, projectionType 0)
// ------------------------------------------------------------------------------------------------
void adaptToFace(const Oracle& o, int nv, const double* Qe /*[b][a][nv]*/, int f, int side, double* dst /*face storage base for nv vars*/) {
    const int n = o.n, N = o.N; Idx ix{n};
    if (side == 0) {
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
            for (int q = 0; q < nv; ++q) dst[ix.fnode(f, 0, i, j) * nv + q] = Qe[(j * n + i) * nv + q];
    } else {
        const int rot = o.faceRot[f];
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
            int ii, jj; leftIndexes2Right(i, j, N, N, rot, ii, jj);
            for (int q = 0; q < nv; ++q) dst[ix.fnode(f, 1, i, j) * nv + q] = Qe[(jj * n + ii) * nv + q];
        }
    }
}

void prolongToFaces(Oracle& o, int nv, const std::vector<double>& field, std::vector<double>& faceField) {
    const int n = o.n; Idx ix{n};
#pragma omp parallel
    {
        std::vector<double> T[6];
        for (int f = 0; f < 6; ++f) T[f].resize((size_t)n * n * nv);
#pragma omp for schedule(static)
        for (int e = 0; e < o.nElem; ++e) {
            for (int f = 0; f < 6; ++f) std::fill(T[f].begin(), T[f].end(), 0.0);
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                const double* q = &field[ix.node(e, i, j, k) * nv];
                for (int c = 0; c < nv; ++c) {
                    T[ELEFT][(k * n + j) * nv + c] = T[ELEFT][(k * n + j) * nv + c] + q[c] * o.v[0 * n + i];
                    T[ERIGHT][(k * n + j) * nv + c] = T[ERIGHT][(k * n + j) * nv + c] + q[c] * o.v[1 * n + i];
                    T[EFRONT][(k * n + i) * nv + c] = T[EFRONT][(k * n + i) * nv + c] + q[c] * o.v[0 * n + j];
                    T[EBACK][(k * n + i) * nv + c] = T[EBACK][(k * n + i) * nv + c] + q[c] * o.v[1 * n + j];
                    T[EBOTTOM][(j * n + i) * nv + c] = T[EBOTTOM][(j * n + i) * nv + c] + q[c] * o.v[0 * n + k];
                    T[ETOP][(j * n + i) * nv + c] = T[ETOP][(j * n + i) * nv + c] + q[c] * o.v[1 * n + k];
                }
            }
            for (int lf = 0; lf < 6; ++lf) adaptToFace(o, nv, T[lf].data(), o.elemFace[6 * e + lf], o.elemFaceSide[6 * e + lf], faceField.data());
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  BR2_ComputeGradient (libs/discretization/EllipticBR2.f90:122-297) and IP_ComputeGradient (EllipticIP.f90:189-362),
//  entered with the local gradients in Ux,Uy,Uz: both prolong the LOCAL gradients first, then lift the interface jumps.
// ------------------------------------------------------------------------------------------------
void computeGradientBR2IP(Oracle& o) {
    const int n = o.n, N = o.N; Idx ix{n};
    const bool br2 = o.ph.viscous == H3D_VISCOUS_BR2;
    prolongToFaces(o, 5, o.Ux, o.fUx); prolongToFaces(o, 5, o.Uy, o.fUy); prolongToFaces(o, 5, o.Uz, o.fUz);
    // BR2_GradientInterfaceSolution / ...Boundary (EllipticBR2.f90:458-592) = IP_GradientInterfaceSolution / ...Boundary
    // (EllipticIP.f90:410-585): Uhat = 1/2 (UL - UR) J_f, unStar_d = Uhat n_d; the boundary state comes from StateForEqn
    // (= FlowState for NCONS equations), not from the gradient-variable BC
#pragma omp parallel for schedule(static)
    for (int f = 0; f < o.nFace; ++f) {
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
            const double Jf = o.fJac[ix.gnode(f, i, j)]; const double* nh = &o.fNormal[3 * ix.gnode(f, i, j)];
            const double* QL = &o.fQ[ix.fnode(f, 0, i, j) * 5];
            double UL[5];
            GetGradients(o, QL, UL);
            double* uL = &o.unStar[ix.fnode(f, 0, i, j) * 15];
            if (o.faceType[f] == H3D_FACE_INTERIOR) {
                double UR[5];
                GetGradients(o, &o.fQ[ix.fnode(f, 1, i, j) * 5], UR);
                int ii, jj; leftIndexes2Right(i, j, N, N, o.faceRot[f], ii, jj);
                double* uR = &o.unStar[ix.fnode(f, 1, ii, jj) * 15];
                for (int q = 0; q < 5; ++q) {
                    double Uhat = 0.5 * (UL[q] - UR[q]) * Jf;
                    for (int d = 0; d < 3; ++d) { double val = Uhat * nh[d]; uL[d * 5 + q] = val; uR[d * 5 + q] = 1 * val; }
                }
            } else if (o.faceType[f] == H3D_FACE_BOUNDARY) {
                double bvExt[5], UR[5];
                for (int q = 0; q < 5; ++q) bvExt[q] = QL[q];
                BC_FlowState(o, o.faceZone[f], nh, bvExt);
                GetGradients(o, bvExt, UR);
                for (int q = 0; q < 5; ++q) {
                    double Uhat = 0.5 * (UL[q] - UR[q]) * Jf;
                    for (int d = 0; d < 3; ++d) uL[d * 5 + q] = Uhat * nh[d];
                }
            }
        }
    }
    // BR2_ComputeGradientFaceIntegrals (EllipticBR2.f90:301-454) / IP_ComputeGradientFaceIntegrals (EllipticIP.f90:366-406)
    const double eta = o.ph.penaltyParameter;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o.nElem; ++e) {
        const double* H[6]; double* FU[6][3];
        for (int lf = 0; lf < 6; ++lf) {
            const size_t fb = ix.fnode(o.elemFace[6 * e + lf], o.elemFaceSide[6 * e + lf], 0, 0);
            H[lf] = &o.unStar[fb * 15];
            FU[lf][0] = &o.fUx[fb * 5]; FU[lf][1] = &o.fUy[fb * 5]; FU[lf][2] = &o.fUz[fb * 5];
        }
        for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
            size_t g = ix.node(e, i, j, k);
            // BR2: invjac = 1/jacobian, U -= faceInt * invjac ; IP: invjac = IPmethod * invJacobian, U += faceInt * invjac
            const double iJ = br2 ? 1.0 / o.jac[g] : o.ph.ipVariant * o.invJac[g];
            for (int d = 0; d < 3; ++d) {
                double* Ud = (d == 0 ? o.Ux.data() : d == 1 ? o.Uy.data() : o.Uz.data()) + 5 * g;
                for (int q = 0; q < 5; ++q) {
                    double fi = H[ELEFT][((k * n + j) * 3 + d) * 5 + q] * o.b[0 * n + i];
                    fi = fi + H[ERIGHT][((k * n + j) * 3 + d) * 5 + q] * o.b[1 * n + i];
                    fi = fi + H[EFRONT][((k * n + i) * 3 + d) * 5 + q] * o.b[0 * n + j];
                    fi = fi + H[EBACK][((k * n + i) * 3 + d) * 5 + q] * o.b[1 * n + j];
                    fi = fi + H[EBOTTOM][((j * n + i) * 3 + d) * 5 + q] * o.b[0 * n + k];
                    fi = fi + H[ETOP][((j * n + i) * 3 + d) * 5 + q] * o.b[1 * n + k];
                    Ud[q] = br2 ? Ud[q] - fi * iJ : Ud[q] + fi * iJ;
                }
            }
        }
        if (!br2) continue;
        // interface gradients correction: the element's own face storage, indexed with the ELEMENT's (j,k) / (i,k) / (i,j)
        // as the reference does (:362-451), whatever the rotation of the face
        const int lfs[6] = {ELEFT, ERIGHT, EFRONT, EBACK, EBOTTOM, ETOP};
        for (int s = 0; s < 6; ++s) {
            const int lf = lfs[s], side = s & 1, ax = s / 2;
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                const int ab = ax == 0 ? k * n + j : (ax == 1 ? k * n + i : j * n + i);
                const int l = ax == 0 ? i : (ax == 1 ? j : k);
                const double bv = o.b[side * n + l] * o.v[side * n + l];
                const double invjac = 1.0 / o.jac[ix.node(e, i, j, k)];
                for (int d = 0; d < 3; ++d) for (int q = 0; q < 5; ++q)
                    FU[lf][d][ab * 5 + q] = FU[lf][d][ab * 5 + q] - eta * H[lf][(ab * 3 + d) * 5 + q] * bv * invjac;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  BR1_ComputeGradient (libs/discretization/EllipticBR1.f90:71-160, 168-393)
// ------------------------------------------------------------------------------------------------
void computeGradient(Oracle& o, double time) {
    (void)time;
    const int n = o.n, n3 = o.n3(); Idx ix{n};
    // HexElement_ComputeLocalGradient (HexElementClass.f90:427-531); U = Q (NSGradientVariables_STATE)
#pragma omp parallel
    {
        std::vector<double> Uxi((size_t)n3 * 5), Ueta((size_t)n3 * 5), Uzeta((size_t)n3 * 5), Ugv((size_t)n3 * 5);
#pragma omp for schedule(static)
        for (int e = 0; e < o.nElem; ++e) {
            std::fill(Uxi.begin(), Uxi.end(), 0.0); std::fill(Ueta.begin(), Ueta.end(), 0.0); std::fill(Uzeta.begin(), Uzeta.end(), 0.0);
            const double* U = &o.Q[ix.node(e, 0, 0, 0) * 5];
            if (o.ph.gradientVariables != H3D_GRADVARS_STATE) {   // GetGradientValues at every node (:466-470)
                for (int t = 0; t < n3; ++t) GetGradients(o, U + (size_t)t * 5, &Ugv[(size_t)t * 5]);
                U = Ugv.data();
            }
            auto L = [&](int i, int j, int k) { return (size_t)((k * n + j) * n + i) * 5; };
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int l = 0; l < n; ++l) for (int i = 0; i < n; ++i)
                for (int q = 0; q < 5; ++q) Uxi[L(i, j, k) + q] = Uxi[L(i, j, k) + q] + U[L(l, j, k) + q] * o.D[i * n + l];
            for (int k = 0; k < n; ++k) for (int l = 0; l < n; ++l) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
                for (int q = 0; q < 5; ++q) Ueta[L(i, j, k) + q] = Ueta[L(i, j, k) + q] + U[L(i, l, k) + q] * o.D[j * n + l];
            for (int l = 0; l < n; ++l) for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
                for (int q = 0; q < 5; ++q) Uzeta[L(i, j, k) + q] = Uzeta[L(i, j, k) + q] + U[L(i, j, l) + q] * o.D[k * n + l];
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                size_t g = ix.node(e, i, j, k);
                const double* jx = &o.JaXi[3 * g]; const double* je = &o.JaEta[3 * g]; const double* jz = &o.JaZeta[3 * g];
                double iJ = o.invJac[g];
                for (int q = 0; q < 5; ++q) {
                    double a = Uxi[L(i, j, k) + q], b = Ueta[L(i, j, k) + q], c = Uzeta[L(i, j, k) + q];
                    o.Ux[5 * g + q] = (a * jx[0] + b * je[0] + c * jz[0]) * iJ;
                    o.Uy[5 * g + q] = (a * jx[1] + b * je[1] + c * jz[1]) * iJ;
                    o.Uz[5 * g + q] = (a * jx[2] + b * je[2] + c * jz[2]) * iJ;
                }
            }
        }
    }
    // Euler with "compute gradients": the viscous discretization is the base class, whose ComputeGradient is the local
    // gradient alone (SpatialDiscretization.f90:185-195, EllipticDiscretizationClass.f90:122-187): no interface terms
    if (!o.ph.flowIsNavierStokes) {
        prolongToFaces(o, 5, o.Ux, o.fUx); prolongToFaces(o, 5, o.Uy, o.fUy); prolongToFaces(o, 5, o.Uz, o.fUz);
        return;
    }
    if (o.ph.viscous != H3D_VISCOUS_BR1) { computeGradientBR2IP(o); return; }
    // BR1_ComputeElementInterfaceAverage (:571-627) + Face_ProjectGradientFluxToElements (FaceClass.f90:865-961, factor = 1)
    // BR1_ComputeBoundaryFlux (:686-736)
#pragma omp parallel for schedule(static)
    for (int f = 0; f < o.nFace; ++f) {
        const int N = o.N;
        if (o.faceType[f] == H3D_FACE_INTERIOR) {
            for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                double UL[5], UR[5];
                GetGradients(o, &o.fQ[ix.fnode(f, 0, i, j) * 5], UL); GetGradients(o, &o.fQ[ix.fnode(f, 1, i, j) * 5], UR);
                const double Jf = o.fJac[ix.gnode(f, i, j)]; const double* nh = &o.fNormal[3 * ix.gnode(f, i, j)];
                int ii, jj; leftIndexes2Right(i, j, N, N, o.faceRot[f], ii, jj);
                double* uL = &o.unStar[ix.fnode(f, 0, i, j) * 15]; double* uR = &o.unStar[ix.fnode(f, 1, ii, jj) * 15];
                for (int q = 0; q < 5; ++q) {
                    double uStar = 0.5 * (UR[q] - UL[q]) * Jf;
                    for (int d = 0; d < 3; ++d) { double val = uStar * nh[d]; uL[d * 5 + q] = val; uR[d * 5 + q] = 1 * val; }
                }
            }
        } else if (o.faceType[f] == H3D_FACE_BOUNDARY) {
            const int zone = o.faceZone[f];
            for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                const double* Qi = &o.fQ[ix.fnode(f, 0, i, j) * 5];
                const double Jf = o.fJac[ix.gnode(f, i, j)]; const double* nh = &o.fNormal[3 * ix.gnode(f, i, j)];
                double u_int[5], u_star[5];
                GetGradients(o, Qi, u_int);
                for (int q = 0; q < 5; ++q) u_star[q] = u_int[q];
                BC_FlowGradVars(o, zone, nh, Qi, u_star);
                double* uL = &o.unStar[ix.fnode(f, 0, i, j) * 15];
                for (int q = 0; q < 5; ++q) for (int d = 0; d < 3; ++d) uL[d * 5 + q] = (u_star[q] - u_int[q]) * nh[d] * Jf;
            }
        }
    }
    // BR1_GradientFaceLoop (:531-569) -> VectorWeakIntegrals_StdFace (DGIntegrals.f90:365-443)
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o.nElem; ++e) {
        const double* H[6];
        for (int lf = 0; lf < 6; ++lf) H[lf] = &o.unStar[ix.fnode(o.elemFace[6 * e + lf], o.elemFaceSide[6 * e + lf], 0, 0) * 15];
        for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
            size_t g = ix.node(e, i, j, k);
            double iJ = o.invJac[g];
            for (int d = 0; d < 3; ++d) {
                double* Ud = (d == 0 ? o.Ux.data() : d == 1 ? o.Uy.data() : o.Uz.data()) + 5 * g;
                for (int q = 0; q < 5; ++q) {
                    double fi = H[ELEFT][((k * n + j) * 3 + d) * 5 + q] * o.b[0 * n + i];
                    fi = fi + H[ERIGHT][((k * n + j) * 3 + d) * 5 + q] * o.b[1 * n + i];
                    fi = fi + H[EFRONT][((k * n + i) * 3 + d) * 5 + q] * o.b[0 * n + j];
                    fi = fi + H[EBACK][((k * n + i) * 3 + d) * 5 + q] * o.b[1 * n + j];
                    fi = fi + H[EBOTTOM][((j * n + i) * 3 + d) * 5 + q] * o.b[0 * n + k];
                    fi = fi + H[ETOP][((j * n + i) * 3 + d) * 5 + q] * o.b[1 * n + k];
                    Ud[q] = Ud[q] + fi * iJ;
                }
            }
        }
    }
    // HexElement_ProlongGradientsToFaces (HexElementClass.f90:304-372)
    prolongToFaces(o, 5, o.Ux, o.fUx);
    prolongToFaces(o, 5, o.Uy, o.fUy);
    prolongToFaces(o, 5, o.Uz, o.fUz);
}

// ------------------------------------------------------------------------------------------------
//  TimeDerivative_ComputeQDot (NavierStokesSolver/SpatialDiscretization.f90:379-685)
// ------------------------------------------------------------------------------------------------
void computeQDot(Oracle& o, double time) {
    (void)time;
    const int n = o.n, n3 = o.n3(), N = o.N; Idx ix{n};
    const bool NS = o.ph.flowIsNavierStokes != 0;
    // viscosity at the elements (:402-412) [+ Smagorinsky :416-436]
    if (NS) {
#pragma omp parallel for schedule(static)
        for (int e = 0; e < o.nElem; ++e) {
            double delta = 0.0;
            if (o.ph.les != H3D_LES_NONE) delta = std::pow(o.volume[e] / (double)(n * n * n), 1.0 / 3.0);
            for (int q = 0; q < n3; ++q) {
                size_t g = (size_t)e * n3 + q;
                get_laminar_mu_kappa(o, &o.Q[5 * g], o.mu[2 * g], o.mu[2 * g + 1]);
                if (o.ph.les != H3D_LES_NONE) {
                    double mut = SmagorinskyViscosity(o, delta, o.dWall.empty() ? 0.0 : o.dWall[g], &o.Q[5 * g], &o.Ux[5 * g], &o.Uy[5 * g], &o.Uz[5 * g]);
                    o.mu[2 * g] = o.mu[2 * g] + mut; o.mu[2 * g + 1] = o.mu[2 * g + 1] + mut * o.ph.mu_to_kappa;
                }
            }
        }
        // compute_viscosity_at_faces (:1345-1398)
#pragma omp parallel for schedule(static)
        for (int f = 0; f < o.nFace; ++f) {
            const int sides = o.faceType[f] == H3D_FACE_INTERIOR ? 2 : 1;
            double delta = 0.0;
            if (o.ph.les != H3D_LES_NONE) delta = std::sqrt(o.fSurface[f] / (double)(n * n));
            for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) for (int s = 0; s < sides; ++s) {
                size_t g = ix.fnode(f, s, i, j);
                get_laminar_mu_kappa(o, &o.fQ[5 * g], o.fmu[2 * g], o.fmu[2 * g + 1]);
                if (o.ph.les != H3D_LES_NONE) {
                    double mut = SmagorinskyViscosity(o, delta, o.fdWall.empty() ? 0.0 : o.fdWall[(size_t)f * n * n + j * n + i], &o.fQ[5 * g], &o.fUx[5 * g], &o.fUy[5 * g], &o.fUz[5 * g]);
                    o.fmu[2 * g] = o.fmu[2 * g] + mut; o.fmu[2 * g + 1] = o.fmu[2 * g + 1] + mut * o.ph.mu_to_kappa;
                }
            }
        }
    }
    // volume integrals: TimeDerivative_VolumetricContribution (:1602-1682)
#pragma omp parallel
    {
        std::vector<double> Finv((size_t)n3 * 15), Fvis((size_t)n3 * 15, 0.0), Fc((size_t)n3 * 15);
        std::vector<double> fS, gS, hS;
        if (o.ph.inviscid == H3D_SPLIT_DG) { fS.resize((size_t)n3 * n * 5); gS.resize((size_t)n3 * n * 5); hS.resize((size_t)n3 * n * 5); }
#pragma omp for schedule(static)
        for (int e = 0; e < o.nElem; ++e) {
            auto L = [&](int i, int j, int k) { return (size_t)((k * n + j) * n + i); };
            // BaseClass_ComputeInnerFluxes (HyperbolicDiscretizationClass.f90:83-152) / BR1_ComputeInnerFluxes (EllipticBR1.f90:740-814)
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                size_t g = ix.node(e, i, j, k), l = L(i, j, k);
                const double* jx = &o.JaXi[3 * g]; const double* je = &o.JaEta[3 * g]; const double* jz = &o.JaZeta[3 * g];
                double F[5][3];
                EulerFlux(o, &o.Q[5 * g], F);
                for (int q = 0; q < 5; ++q) {
                    Finv[(l * 3 + 0) * 5 + q] = F[q][IX] * jx[IX] + F[q][IY] * jx[IY] + F[q][IZ] * jx[IZ];
                    Finv[(l * 3 + 1) * 5 + q] = F[q][IX] * je[IX] + F[q][IY] * je[IY] + F[q][IZ] * je[IZ];
                    Finv[(l * 3 + 2) * 5 + q] = F[q][IX] * jz[IX] + F[q][IY] * jz[IY] + F[q][IZ] * jz[IZ];
                }
                if (NS) {
                    ViscousFlux(o, &o.Q[5 * g], &o.Ux[5 * g], &o.Uy[5 * g], &o.Uz[5 * g], o.mu[2 * g], 0.0, o.mu[2 * g + 1], F);
                    for (int q = 0; q < 5; ++q) {
                        Fvis[(l * 3 + 0) * 5 + q] = F[q][IX] * jx[IX] + F[q][IY] * jx[IY] + F[q][IZ] * jx[IZ];
                        Fvis[(l * 3 + 1) * 5 + q] = F[q][IX] * je[IX] + F[q][IY] * je[IY] + F[q][IZ] * je[IZ];
                        Fvis[(l * 3 + 2) * 5 + q] = F[q][IX] * jz[IX] + F[q][IY] * jz[IY] + F[q][IZ] * jz[IZ];
                    }
                }
            }
            double* qd = &o.QDot[ix.node(e, 0, 0, 0) * 5];
            if (o.ph.inviscid == H3D_STANDARD_DG) {
                // contravariantFlux = inviscid - viscous - Avisc(=0) ; ScalarWeakIntegrals_StdVolumeGreen (DGIntegrals.f90:56-87)
                for (size_t t = 0; t < (size_t)n3 * 15; ++t) Fc[t] = Finv[t] - Fvis[t] - 0.0;
                for (size_t t = 0; t < (size_t)n3 * 5; ++t) qd[t] = 0.0;
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int l = 0; l < n; ++l) for (int i = 0; i < n; ++i)
                    for (int q = 0; q < 5; ++q) qd[L(i, j, k) * 5 + q] = qd[L(i, j, k) * 5 + q] + o.hatD[i * n + l] * Fc[(L(l, j, k) * 3 + 0) * 5 + q];
                for (int k = 0; k < n; ++k) for (int l = 0; l < n; ++l) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
                    for (int q = 0; q < 5; ++q) qd[L(i, j, k) * 5 + q] = qd[L(i, j, k) * 5 + q] + o.hatD[j * n + l] * Fc[(L(i, l, k) * 3 + 1) * 5 + q];
                for (int l = 0; l < n; ++l) for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
                    for (int q = 0; q < 5; ++q) qd[L(i, j, k) * 5 + q] = qd[L(i, j, k) * 5 + q] + o.hatD[k * n + l] * Fc[(L(i, j, l) * 3 + 2) * 5 + q];
            } else {
                // SplitDG_ComputeSplitFormFluxes (HyperbolicSplitForm.f90:64-116); fSharp(:,l,i,j,k) -> fS[(l + n*node)*5 + q]
                auto SI = [&](int l, int i, int j, int k) { return ((size_t)L(i, j, k) * n + l) * 5; };
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) for (int q = 0; q < 5; ++q) {
                    fS[SI(i, i, j, k) + q] = Finv[(L(i, j, k) * 3 + 0) * 5 + q];
                    gS[SI(j, i, j, k) + q] = Finv[(L(i, j, k) * 3 + 1) * 5 + q];
                    hS[SI(k, i, j, k) + q] = Finv[(L(i, j, k) * 3 + 2) * 5 + q];
                }
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                    size_t g = ix.node(e, i, j, k);
                    for (int l = i + 1; l < n; ++l) {
                        size_t g2 = ix.node(e, l, j, k);
                        TwoPointFlux(o, &o.Q[5 * g], &o.Q[5 * g2], &o.JaXi[3 * g], &o.JaXi[3 * g2], &fS[SI(l, i, j, k)]);
                        for (int q = 0; q < 5; ++q) fS[SI(i, l, j, k) + q] = fS[SI(l, i, j, k) + q];
                    }
                    for (int l = j + 1; l < n; ++l) {
                        size_t g2 = ix.node(e, i, l, k);
                        TwoPointFlux(o, &o.Q[5 * g], &o.Q[5 * g2], &o.JaEta[3 * g], &o.JaEta[3 * g2], &gS[SI(l, i, j, k)]);
                        for (int q = 0; q < 5; ++q) gS[SI(j, i, l, k) + q] = gS[SI(l, i, j, k) + q];
                    }
                    for (int l = k + 1; l < n; ++l) {
                        size_t g2 = ix.node(e, i, j, l);
                        TwoPointFlux(o, &o.Q[5 * g], &o.Q[5 * g2], &o.JaZeta[3 * g], &o.JaZeta[3 * g2], &hS[SI(l, i, j, k)]);
                        for (int q = 0; q < 5; ++q) hS[SI(k, i, j, l) + q] = hS[SI(l, i, j, k) + q];
                    }
                }
                // ScalarWeakIntegrals_SplitVolumeDivergence (DGIntegrals.f90:92-129); QDot = -volInt (:1678)
                for (size_t t = 0; t < (size_t)n3 * 5; ++t) qd[t] = 0.0;
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int l = 0; l < n; ++l) for (int i = 0; i < n; ++i)
                    for (int q = 0; q < 5; ++q) qd[L(i, j, k) * 5 + q] = qd[L(i, j, k) * 5 + q] + o.sharpD[i * n + l] * fS[SI(l, i, j, k) + q] + o.hatD[i * n + l] * Fvis[(L(l, j, k) * 3 + 0) * 5 + q];
                for (int k = 0; k < n; ++k) for (int l = 0; l < n; ++l) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
                    for (int q = 0; q < 5; ++q) qd[L(i, j, k) * 5 + q] = qd[L(i, j, k) * 5 + q] + o.sharpD[j * n + l] * gS[SI(l, i, j, k) + q] + o.hatD[j * n + l] * Fvis[(L(i, l, k) * 3 + 1) * 5 + q];
                for (int l = 0; l < n; ++l) for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
                    for (int q = 0; q < 5; ++q) qd[L(i, j, k) * 5 + q] = qd[L(i, j, k) * 5 + q] + o.sharpD[k * n + l] * hS[SI(l, i, j, k) + q] + o.hatD[k * n + l] * Fvis[(L(i, j, l) * 3 + 2) * 5 + q];
                for (size_t t = 0; t < (size_t)n3 * 5; ++t) qd[t] = -qd[t];
            }
        }
    }
    // Riemann solver of non-shared faces: computeElementInterfaceFlux (:1710-1799), computeBoundaryFlux (:1896-2028)
#pragma omp parallel for schedule(static)
    for (int f = 0; f < o.nFace; ++f) {
        if (o.faceType[f] == H3D_FACE_INTERIOR) {
            for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                size_t gL = ix.fnode(f, 0, i, j), gR = ix.fnode(f, 1, i, j), gg = ix.gnode(f, i, j);
                const double* nh = &o.fNormal[3 * gg];
                double visc[5] = {0, 0, 0, 0, 0}, inv[5];
                if (NS) {   // BR1_RiemannSolver (EllipticBR1.f90:816-868)
                    double fL[5][3], fR[5][3];
                    ViscousFlux(o, &o.fQ[5 * gL], &o.fUx[5 * gL], &o.fUy[5 * gL], &o.fUz[5 * gL], o.fmu[2 * gL], 0.0, o.fmu[2 * gL + 1], fL);
                    ViscousFlux(o, &o.fQ[5 * gR], &o.fUx[5 * gR], &o.fUy[5 * gR], &o.fUz[5 * gR], o.fmu[2 * gR], 0.0, o.fmu[2 * gR + 1], fR);
                    for (int q = 0; q < 5; ++q) {
                        double fx = 0.5 * (fL[q][IX] + fR[q][IX]), fy = 0.5 * (fL[q][IY] + fR[q][IY]), fz = 0.5 * (fL[q][IZ] + fR[q][IZ]);
                        visc[q] = fx * nh[IX] + fy * nh[IY] + fz * nh[IZ];
                        // IP_RiemannSolver (EllipticIP.f90:704-761) with PenaltyParameterNS (:678-687)
                        if (o.ph.viscous == H3D_VISCOUS_IP) {
                            const double penalty = 0.5 * o.ph.penaltyParameter * (N + 1) * (N + 2) / o.fH[f];
                            visc[q] = visc[q] - penalty * o.ph.mu * (o.fQ[5 * gL + q] - o.fQ[5 * gR + q]);
                        }
                    }
                }
                RiemannSolver(o, &o.fQ[5 * gL], &o.fQ[5 * gR], nh, &o.fT1[3 * gg], &o.fT2[3 * gg], inv);
                // Face_ProjectFluxToElements (FaceClass.f90:597-696): left = flux, right = rotated and negated
                int ii, jj; leftIndexes2Right(i, j, N, N, o.faceRot[f], ii, jj);
                size_t gRe = ix.fnode(f, 1, ii, jj);
                for (int q = 0; q < 5; ++q) {
                    double flux = (inv[q] - visc[q]) * o.fJac[gg] - 0.0;
                    o.fStar[5 * gL + q] = flux; o.fStar[5 * gRe + q] = -flux;
                }
            }
        } else if (o.faceType[f] == H3D_FACE_BOUNDARY) {
            const int zone = o.faceZone[f];
            for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                size_t gL = ix.fnode(f, 0, i, j), gR = ix.fnode(f, 1, i, j), gg = ix.gnode(f, i, j);
                const double* nh = &o.fNormal[3 * gg];
                for (int q = 0; q < 5; ++q) o.fQ[5 * gR + q] = o.fQ[5 * gL + q];
                BC_FlowState(o, zone, nh, &o.fQ[5 * gR]);
                double visc[5] = {0, 0, 0, 0, 0}, inv[5];
                if (NS) {
                    double fv[5][3];
                    ViscousFlux(o, &o.fQ[5 * gL], &o.fUx[5 * gL], &o.fUy[5 * gL], &o.fUz[5 * gL], o.fmu[2 * gL], 0.0, o.fmu[2 * gL + 1], fv);
                    for (int q = 0; q < 5; ++q) { visc[q] = fv[q][IX] * nh[IX] + fv[q][IY] * nh[IY] + fv[q][IZ] * nh[IZ]; visc[q] = visc[q] + 0.0; }
                    BC_FlowNeumann(o, zone, &o.fQ[5 * gL], visc);
                }
                RiemannSolver(o, &o.fQ[5 * gL], &o.fQ[5 * gR], nh, &o.fT1[3 * gg], &o.fT2[3 * gg], inv);
                for (int q = 0; q < 5; ++q) o.fStar[5 * gL + q] = (inv[q] - visc[q]) * o.fJac[gg];
            }
        }
    }
    // surface integrals + scaling: TimeDerivative_FacesContribution (:1686-1701) -> ScalarWeakIntegrals_StdFace
    // (DGIntegrals.f90:214-273); QDot /= jacobian (:489-491); QDot += S_NS (:632-638)
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o.nElem; ++e) {
        const double* F[6];
        for (int lf = 0; lf < 6; ++lf) F[lf] = &o.fStar[ix.fnode(o.elemFace[6 * e + lf], o.elemFaceSide[6 * e + lf], 0, 0) * 5];
        for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
            size_t g = ix.node(e, i, j, k);
            for (int q = 0; q < 5; ++q) {
                double fi = F[ELEFT][(k * n + j) * 5 + q] * o.b[0 * n + i];
                fi = fi + F[ERIGHT][(k * n + j) * 5 + q] * o.b[1 * n + i];
                fi = fi + F[EFRONT][(k * n + i) * 5 + q] * o.b[0 * n + j];
                fi = fi + F[EBACK][(k * n + i) * 5 + q] * o.b[1 * n + j];
                fi = fi + F[EBOTTOM][(j * n + i) * 5 + q] * o.b[0 * n + k];
                fi = fi + F[ETOP][(j * n + i) * 5 + q] * o.b[1 * n + k];
                double r = o.QDot[5 * g + q] - fi;
                r = r / o.jac[g];
                r = r + (o.hasSource ? o.S[5 * g + q] : 0.0);
                o.QDot[5 * g + q] = r;
            }
        }
    }
}

// ComputeTimeDerivative (SpatialDiscretization.f90:227-320)
void computeTimeDerivativeP(Oracle& o, double time);   // h3d_oracle_p.inc
void computeTimeDerivative(Oracle& o, double time) {
    if (o.mixed) { computeTimeDerivativeP(o, time); return; }
    prolongToFaces(o, 5, o.Q, o.fQ);
    if (o.ph.computeGradients) computeGradient(o, time);
    computeQDot(o, time);
}

}  // namespace

#include "h3d_oracle_p.inc"

// ====================================================================================================
//  C API (mirrors include/h3d_gpu.h with the prefix orc_)
// ====================================================================================================
extern "C" {

// threads the OpenMP loops of the oracle really use (bench.py reports this as cpu_baseline.cores), and a setter for launchers
// that pin OMP_NUM_THREADS=1 before the process starts (torchrun)
int orc_num_threads() {
#ifdef _OPENMP
    int nt = 1;
#pragma omp parallel
    {
#pragma omp single
        nt = omp_get_num_threads();
    }
    return nt;
#else
    return 1;
#endif
}
void orc_set_num_threads(int nt) {
#ifdef _OPENMP
    if (nt > 0) omp_set_num_threads(nt);
#else
    (void)nt;
#endif
}

void* orc_create() { return new Oracle(); }
int orc_create_handle(void** out, int, int, int, const void*) { *out = new Oracle(); return 0; }   // the signature of h3d_create
void orc_destroy(void* p) { Oracle* o = (Oracle*)p; delete (PData*)o->pdata; delete o; }
const char* orc_last_error(void* p) { return ((Oracle*)p)->err.c_str(); }

int orc_set_physics(void* p, const H3dPhysics* ph) { ((Oracle*)p)->ph = *ph; return 0; }

int orc_set_basis(void* p, int N, int nodeType, const double* x, const double* w, const double* D, const double* hatD,
                  const double* sharpD, const double* v, const double* b) {
    Oracle& o = *(Oracle*)p; const int n = N + 1;
    o.N = N; o.n = n; o.nodeType = nodeType;
    o.x.assign(x, x + n); o.w.assign(w, w + n); o.D.assign(D, D + n * n); o.hatD.assign(hatD, hatD + n * n);
    o.sharpD.assign(sharpD, sharpD + n * n); o.v.assign(v, v + 2 * n); o.b.assign(b, b + 2 * n);
    // every order set so far stays registered: NodalStorage(N) of a p-nonconforming mesh
    if (!o.pdata) o.pdata = new PData();
    Basis1& bs = PD(o).sp[N];
    bs.N = N; bs.n = n; bs.nodeType = nodeType; bs.x = o.x; bs.w = o.w; bs.D = o.D; bs.hatD = o.hatD; bs.sharpD = o.sharpD; bs.v = o.v; bs.b = o.b;
    return 0;
}

// Tset(Norigin, Ndest) % T (libs/spectral/InterpolationMatrices.f90:42-107), row-major [(Ndest+1)][(Norigin+1)]
int orc_set_interpolation(void* p, int Norigin, int Ndest, const double* T) {
    Oracle& o = *(Oracle*)p;
    if (!o.pdata) o.pdata = new PData();
    PD(o).T[{Norigin, Ndest}].assign(T, T + (size_t)(Norigin + 1) * (Ndest + 1));
    return 0;
}

// h3d_set_mesh_p: as orc_set_mesh with the elements' orders elemOrder[nElem][3]; geometry arrays packed element after element /
// face after face at their own sizes (faces at the face order, FaceClass.f90:187-282)
int orc_set_mesh_p(void* p, int nElem, int nFace, const int* elemOrder, const int* faceOrder, const int* elemFace, const int* elemFaceSide, const int* faceElem,
                   const int* faceElemSide, const int* faceRot, const int* faceType, const int* faceZone,
                   const double* jGradXi, const double* jGradEta, const double* jGradZeta, const double* jacobian,
                   const double* x, const double* volume, const double* faceNormal, const double* faceT1, const double* faceT2,
                   const double* faceJacobian, const double* faceX, const double* faceSurface) {
    Oracle& o = *(Oracle*)p;
    if (!o.pdata) { o.err = "set_basis must precede set_mesh_p"; return 1; }
    if (o.ph.flowIsNavierStokes && o.ph.viscous == H3D_VISCOUS_BR2) { o.err = "p-nonconforming meshes: BR1 or interior penalty (or Euler)"; return 1; }
    if (o.ph.les != H3D_LES_NONE && (!volume || !faceSurface)) { o.err = "LES needs the element volumes and face surfaces"; return 1; }
    PData& P = PD(o);
    if (o.ph.inviscid == H3D_SPLIT_DG) for (int q = 0; q < 3 * nElem; ++q) {
        auto it = P.sp.find(elemOrder[q]);
        if (it != P.sp.end() && it->second.nodeType != H3D_GAUSSLOBATTO) { o.err = "split-form discretization needs Gauss-Lobatto nodes"; return 1; }
    }
    o.mixed = true; o.nElem = nElem; o.nFace = nFace;
    o.elemFace.assign(elemFace, elemFace + 6 * (size_t)nElem); o.elemFaceSide.assign(elemFaceSide, elemFaceSide + 6 * (size_t)nElem);
    o.faceElem.assign(faceElem, faceElem + 2 * (size_t)nFace); o.faceElemSide.assign(faceElemSide, faceElemSide + 2 * (size_t)nFace);
    o.faceRot.assign(faceRot, faceRot + nFace); o.faceType.assign(faceType, faceType + nFace); o.faceZone.assign(faceZone, faceZone + nFace);
    for (int f = 0; f < nFace; ++f) if (faceType[f] == H3D_FACE_MPI) { o.err = "the oracle is single-domain: MPI faces are not supported"; return 1; }
    P.Nxyz.assign(elemOrder, elemOrder + 3 * (size_t)nElem);
    P.eOff.assign(nElem + 1, 0); P.tOff.assign(6 * (size_t)nElem + 1, 0);
    for (int e = 0; e < nElem; ++e) {
        for (int d = 0; d < 3; ++d) if (!P.sp.count(elemOrder[3 * e + d])) { o.err = "set_basis has not been called for every polynomial order of the mesh"; return 1; }
        P.eOff[e + 1] = P.eOff[e] + (size_t)(elemOrder[3 * e] + 1) * (elemOrder[3 * e + 1] + 1) * (elemOrder[3 * e + 2] + 1);
        for (int lf = 0; lf < 6; ++lf) { int Nel[2]; elemFaceOrders(P, e, lf, Nel); P.tOff[6 * e + lf + 1] = P.tOff[6 * e + lf] + (size_t)(Nel[0] + 1) * (Nel[1] + 1); }
    }
    // Face_LinkWithElements (FaceClass.f90:187-282)
    P.fo.assign(6 * (size_t)nFace, 0); P.proj.assign(2 * (size_t)nFace, 0); P.fOff.assign(nFace + 1, 0);
    for (int f = 0; f < nFace; ++f) {
        int NelL[2], NelR[2];
        elemFaceOrders(P, faceElem[2 * f], faceElemSide[2 * f], NelL);
        NelR[0] = NelL[0]; NelR[1] = NelL[1];
        if (faceType[f] == H3D_FACE_INTERIOR) elemFaceOrders(P, faceElem[2 * f + 1], faceElemSide[2 * f + 1], NelR);
        int NfR[2] = {NelR[0], NelR[1]};
        const int rot = faceRot[f];
        if (rot == 1 || rot == 3 || rot == 4 || rot == 6) { NfR[0] = NelR[1]; NfR[1] = NelR[0]; }
        int* fo = &P.fo[6 * (size_t)f];
        fo[0] = std::max(NelL[0], NfR[0]); fo[1] = std::max(NelL[1], NfR[1]); fo[2] = NelL[0]; fo[3] = NelL[1]; fo[4] = NfR[0]; fo[5] = NfR[1];
        for (int s = 0; s < 2; ++s) {
            P.proj[2 * f + s] = (fo[2 + 2 * s] != fo[0] ? 1 : 0) + (fo[3 + 2 * s] != fo[1] ? 2 : 0);
            for (int d = 0; d < 2; ++d) if (fo[2 + 2 * s + d] != fo[d] && (!P.T.count({fo[2 + 2 * s + d], fo[d]}) || !P.T.count({fo[d], fo[2 + 2 * s + d]}))) {
                o.err = "set_interpolation has not been called for every pair of orders that meet at a face"; return 1; }
        }
        if (!P.sp.count(fo[0]) || !P.sp.count(fo[1])) { o.err = "set_basis has not been called for every face order"; return 1; }
        if (faceOrder) for (int q = 0; q < 6; ++q) if (faceOrder[6 * (size_t)f + q] != fo[q]) { o.err = "faceOrder contradicts the orders of the elements"; return 1; }
        P.fOff[f + 1] = P.fOff[f] + (size_t)(fo[0] + 1) * (fo[1] + 1);
    }
    const size_t ne = P.eOff[nElem], nfn = P.fOff[nFace], nt = P.tOff[6 * (size_t)nElem];
    o.JaXi.assign(jGradXi, jGradXi + 3 * ne); o.JaEta.assign(jGradEta, jGradEta + 3 * ne); o.JaZeta.assign(jGradZeta, jGradZeta + 3 * ne);
    o.jac.assign(jacobian, jacobian + ne); o.invJac.resize(ne);
    for (size_t q = 0; q < ne; ++q) o.invJac[q] = 1.0 / o.jac[q];
    if (x) o.xyz.assign(x, x + 3 * ne);
    if (volume) o.volume.assign(volume, volume + nElem);
    o.fNormal.assign(faceNormal, faceNormal + 3 * nfn); o.fT1.assign(faceT1, faceT1 + 3 * nfn); o.fT2.assign(faceT2, faceT2 + 3 * nfn);
    o.fJac.assign(faceJacobian, faceJacobian + nfn);
    if (faceX) o.fX.assign(faceX, faceX + 3 * nfn);
    if (faceSurface) o.fSurface.assign(faceSurface, faceSurface + nFace);
    o.Q.assign(5 * ne, 0.0); o.QDot.assign(5 * ne, 0.0); o.G.assign(5 * ne, 0.0); o.S.assign(5 * ne, 0.0);
    o.Ux.assign(5 * ne, 0.0); o.Uy.assign(5 * ne, 0.0); o.Uz.assign(5 * ne, 0.0); o.mu.assign(2 * ne, 0.0);
    o.fQ.assign(10 * nfn, 0.0); o.fUx.assign(10 * nfn, 0.0); o.fUy.assign(10 * nfn, 0.0); o.fUz.assign(10 * nfn, 0.0); o.fmu.assign(4 * nfn, 0.0);
    P.fStarE.assign(5 * nt, 0.0); P.unStarE.assign(15 * nt, 0.0);
    return 0;
}

int orc_set_mesh(void* p, int nElem, int nFace, const int* elemFace, const int* elemFaceSide, const int* faceElem,
                 const int* faceElemSide, const int* faceRot, const int* faceType, const int* faceZone,
                 const double* jGradXi, const double* jGradEta, const double* jGradZeta, const double* jacobian,
                 const double* x, const double* volume, const double* faceNormal, const double* faceT1, const double* faceT2,
                 const double* faceJacobian, const double* faceX, const double* faceSurface) {
    Oracle& o = *(Oracle*)p;
    if (o.N < 0) { o.err = "set_basis must precede set_mesh"; return 1; }
    const size_t n3 = o.n3(), n2 = o.n2();
    o.nElem = nElem; o.nFace = nFace;
    o.elemFace.assign(elemFace, elemFace + 6 * (size_t)nElem); o.elemFaceSide.assign(elemFaceSide, elemFaceSide + 6 * (size_t)nElem);
    o.faceElem.assign(faceElem, faceElem + 2 * (size_t)nFace); o.faceElemSide.assign(faceElemSide, faceElemSide + 2 * (size_t)nFace);
    o.faceRot.assign(faceRot, faceRot + nFace); o.faceType.assign(faceType, faceType + nFace); o.faceZone.assign(faceZone, faceZone + nFace);
    for (int f = 0; f < nFace; ++f) if (faceType[f] == H3D_FACE_MPI) { o.err = "the oracle is single-domain: MPI faces are not supported"; return 1; }
    o.JaXi.assign(jGradXi, jGradXi + 3 * n3 * nElem); o.JaEta.assign(jGradEta, jGradEta + 3 * n3 * nElem); o.JaZeta.assign(jGradZeta, jGradZeta + 3 * n3 * nElem);
    o.jac.assign(jacobian, jacobian + n3 * nElem); o.invJac.resize(n3 * nElem);
    for (size_t q = 0; q < n3 * nElem; ++q) o.invJac[q] = 1.0 / o.jac[q];   // MappedGeometry.f90:382
    if (x) o.xyz.assign(x, x + 3 * n3 * nElem);
    if (volume) o.volume.assign(volume, volume + nElem);
    o.fNormal.assign(faceNormal, faceNormal + 3 * n2 * nFace); o.fT1.assign(faceT1, faceT1 + 3 * n2 * nFace); o.fT2.assign(faceT2, faceT2 + 3 * n2 * nFace);
    o.fJac.assign(faceJacobian, faceJacobian + n2 * nFace);
    if (faceX) o.fX.assign(faceX, faceX + 3 * n2 * nFace);
    if (faceSurface) o.fSurface.assign(faceSurface, faceSurface + nFace);
    const size_t ne = n3 * nElem, nf = n2 * nFace * 2;
    o.Q.assign(5 * ne, 0.0); o.QDot.assign(5 * ne, 0.0); o.G.assign(5 * ne, 0.0); o.S.assign(5 * ne, 0.0);
    o.Ux.assign(5 * ne, 0.0); o.Uy.assign(5 * ne, 0.0); o.Uz.assign(5 * ne, 0.0); o.mu.assign(2 * ne, 0.0);
    o.fQ.assign(5 * nf, 0.0); o.fUx.assign(5 * nf, 0.0); o.fUy.assign(5 * nf, 0.0); o.fUz.assign(5 * nf, 0.0);
    o.fStar.assign(5 * nf, 0.0); o.fmu.assign(2 * nf, 0.0); o.unStar.assign(15 * nf, 0.0);
    return 0;
}

int orc_set_boundary_conditions(void* p, int nZones, const int* bcType, const double* bcParams) {
    Oracle& o = *(Oracle*)p;
    o.nZones = nZones; o.bcType.assign(bcType, bcType + nZones); o.bcParams.assign(bcParams, bcParams + 16 * (size_t)nZones);
    return 0;
}

int orc_set_wall_distance(void* p, const double* dWallElem, const double* dWallFace) {
    Oracle& o = *(Oracle*)p;
    if (o.mixed) { o.dWall.assign(dWallElem, dWallElem + PD(o).eOff[o.nElem]); o.fdWall.assign(dWallFace, dWallFace + PD(o).fOff[o.nFace]); return 0; }
    o.dWall.assign(dWallElem, dWallElem + (size_t)o.nElem * o.n3()); o.fdWall.assign(dWallFace, dWallFace + (size_t)o.nFace * o.n * o.n);
    return 0;
}

int orc_set_face_h(void* p, const double* faceH) { Oracle& o = *(Oracle*)p; o.fH.assign(faceH, faceH + o.nFace); return 0; }

int orc_upload_Q(void* p, const double* Q) { Oracle& o = *(Oracle*)p; std::memcpy(o.Q.data(), Q, o.Q.size() * sizeof(double)); return 0; }

int orc_download(void* p, double* Q, double* QDot, double* Ux, double* Uy, double* Uz) {
    Oracle& o = *(Oracle*)p; const size_t b = o.Q.size() * sizeof(double);
    if (Q) std::memcpy(Q, o.Q.data(), b);
    if (QDot) std::memcpy(QDot, o.QDot.data(), b);
    if (Ux) std::memcpy(Ux, o.Ux.data(), b);
    if (Uy) std::memcpy(Uy, o.Uy.data(), b);
    if (Uz) std::memcpy(Uz, o.Uz.data(), b);
    return 0;
}

// face traces of the last residual evaluation, for unit parity of the prolongation: [f][side][j][i][5]
int orc_download_faces(void* p, double* fQ, double* fUx, double* fUy, double* fUz, double* fStar) {
    if (((Oracle*)p)->mixed) { ((Oracle*)p)->err = "download_faces is not available on p-nonconforming meshes"; return 1; }
    Oracle& o = *(Oracle*)p; const size_t b = o.fQ.size() * sizeof(double);
    if (fQ) std::memcpy(fQ, o.fQ.data(), b);
    if (fUx) std::memcpy(fUx, o.fUx.data(), b);
    if (fUy) std::memcpy(fUy, o.fUy.data(), b);
    if (fUz) std::memcpy(fUz, o.fUz.data(), b);
    if (fStar) std::memcpy(fStar, o.fStar.data(), b);
    return 0;
}

int orc_set_source(void* p, const double* S) {
    Oracle& o = *(Oracle*)p;
    o.hasSource = S != nullptr;
    if (S) std::memcpy(o.S.data(), S, o.S.size() * sizeof(double));
    return 0;
}

int orc_compute_time_derivative(void* p, double time) { computeTimeDerivative(*(Oracle*)p, time); return 0; }

// TakeRK3Step / TakeRK5Step (libs/timeintegrator/ExplicitMethods.f90:667-788, 790-882)
namespace {
// libs/timeintegrator/ExplicitMethods.f90: Euler :1232-1284, RK3 :690-692, RK5 :812-816, LSERK14-4 :903-905, SSPRK33 :999-1002, SSPRK43 :1125-1128
const double RK_A1[1] = {0.0}, RK_B1[1] = {0.0}, RK_C1[1] = {1.0};
const double RK_A3[3] = {0.0, -5.0 / 9.0, -153.0 / 128.0}, RK_B3[3] = {0.0, 1.0 / 3.0, 3.0 / 4.0}, RK_C3[3] = {1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0};
const double RK_A5[5] = {0.0, -0.4178904745, -1.192151694643, -1.697784692471, -1.514183444257};
const double RK_B5[5] = {0.0, 0.1496590219993, 0.3704009573644, 0.6222557631345, 0.9582821306748};
const double RK_C5[5] = {0.1496590219993, 0.3792103129999, 0.8229550293869, 0.6994504559488, 0.1530572479681};
const double RK_A14[14] = {0.0000000000000000, -0.7188012108672410, -0.7785331173421570, -0.0053282796654044, -0.8552979934029281, -3.9564138245774565, -1.5780575380587385,
                           -2.0837094552574054, -0.7483334182761610, -0.7032861106563359, +0.0013917096117681, -0.0932075369637460, -0.9514200470875948, -7.1151571693922548};
const double RK_B14[14] = {0.0000000000000000, 0.0367762454319673, 0.1249685262725025, 0.2446177702277698, 0.2476149531070420, 0.2969311120382472, 0.3978149645802642,
                           0.5270854589440328, 0.6981269994175695, 0.8190890835352128, 0.8527059887098624, 0.8604711817462826, 0.8627060376969976, 0.8734213127600976};
const double RK_C14[14] = {0.0367762454319673, 0.3136296607553959, 0.1531848691869027, 0.0030097086818182, 0.3326293790646110, 0.2440251405350864, 0.3718879239592277,
                           0.6204126221582444, 0.1524043173028741, 0.0760894927419266, 0.0077604214040978, 0.0024647284755382, 0.0780348340049386, 5.5059777270269628};
const double SSP33_A[3] = {1.0, 3.0 / 4.0, 1.0 / 3.0}, SSP33_B[3] = {0.0, 1.0 / 4.0, 2.0 / 3.0}, SSP33_C[3] = {1.0, 1.0 / 4.0, 2.0 / 3.0}, SSP33_D[3] = {0.0, 1.0, 0.5};
const double SSP43_A[4] = {1.0, 0.0, 2.0 / 3.0, 0.0}, SSP43_B[4] = {0.0, 1.0, 1.0 / 3.0, 1.0}, SSP43_C[4] = {0.5, 0.5, 1.0 / 6.0, 0.5}, SSP43_D[4] = {0.0, 0.5, 1.0, 0.5};
int rkStages(int scheme) {
    switch (scheme) {
        case H3D_EULER: return 1; case H3D_RK3: return 3; case H3D_RK5: return 5; case H3D_LSERK14_4: return 14;
        case H3D_SSPRK33: return 3; case H3D_SSPRK43: return 4; default: return 0;
    }
}
// stage_limiter (ExplicitMethods.f90:1755-1847)
void stageLimiter(Oracle& o) {
    Idx ix{o.n};
    const double gm1 = o.ph.gammaMinus1, LIMITER_MIN = o.limiterMin;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o.nElem; ++e) {
        // p-nonconforming meshes: the element's own nodal storages (ExplicitMethods.f90:1776-1779)
        const int nx = o.mixed ? PD(o).Nxyz[3 * e] + 1 : o.n, ny = o.mixed ? PD(o).Nxyz[3 * e + 1] + 1 : o.n, nz = o.mixed ? PD(o).Nxyz[3 * e + 2] + 1 : o.n;
        const double* wx = o.mixed ? PD(o).sp.at(nx - 1).w.data() : o.w.data(); const double* wy = o.mixed ? PD(o).sp.at(ny - 1).w.data() : o.w.data();
        const double* wz = o.mixed ? PD(o).sp.at(nz - 1).w.data() : o.w.data();
        const size_t g0 = o.mixed ? PD(o).eOff[e] : ix.node(e, 0, 0, 0);
        const int n3 = nx * ny * nz;
        double Qavg[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            size_t g = g0 + (size_t)(k * ny + j) * nx + i;
            for (int q = 0; q < 5; ++q) Qavg[q] = Qavg[q] + o.Q[5 * g + q] * wx[i] * wy[j] * wz[k] * o.jac[g];
        }
        for (int q = 0; q < 5; ++q) Qavg[q] = Qavg[q] / o.volume[e];
        double minrho = std::numeric_limits<double>::max();
        for (int t = 0; t < n3; ++t) { double rho = o.Q[5 * (g0 + t)]; if (rho < minrho) minrho = rho; }
        if (Qavg[0] != minrho) {
            double m = std::fmin(LIMITER_MIN, Qavg[0]);
            double theta = std::fabs((Qavg[0] - m) / (Qavg[0] - minrho));
            if (theta <= 1.0)
                for (int t = 0; t < n3; ++t) { double& r = o.Q[5 * (g0 + t)]; r = theta * (r - Qavg[0]) + Qavg[0]; }
        }
        double minp = std::numeric_limits<double>::max(), pavg = 0.0;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            size_t g = g0 + (size_t)(k * ny + j) * nx + i;
            const double* Q = &o.Q[5 * g];
            double p = gm1 * (Q[4] - 0.5 * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]) / Q[0]);
            pavg = pavg + p * wx[i] * wy[j] * wz[k] * o.jac[g];
            if (p < minp) minp = p;
        }
        pavg = pavg / o.volume[e];
        if (pavg != minp) {
            double m = std::fmin(LIMITER_MIN, pavg);
            double theta = std::fabs((pavg - m) / (pavg - minp));
            if (theta <= 1.0)
                for (int t = 0; t < n3; ++t) { double* Q = &o.Q[5 * (g0 + t)]; for (int q = 0; q < 5; ++q) Q[q] = theta * (Q[q] - Qavg[q]) + Qavg[q]; }
        }
    }
}

double rkStage(Oracle& o, int scheme, int k, double t, double dt) {   // loop bodies of the steppers
    if (scheme == H3D_SSPRK33 || scheme == H3D_SSPRK43) {
        const bool s3 = scheme == H3D_SSPRK33;
        const double *a = s3 ? SSP33_A : SSP43_A, *b = s3 ? SSP33_B : SSP43_B, *c = s3 ? SSP33_C : SSP43_C, *d = s3 ? SSP33_D : SSP43_D;
        if (k == 0) o.G = o.Q;
        const double tk = t + d[k] * dt;
        computeTimeDerivative(o, tk);
        const double ak = a[k], bk = b[k], ck = c[k];
#pragma omp parallel for schedule(static)
        for (size_t q = 0; q < o.Q.size(); ++q) o.Q[q] = ak * o.G[q] + bk * o.Q[q] + ck * dt * o.QDot[q];
        if (o.limited) stageLimiter(o);     // :1050-1052, :1177-1179
        return tk;
    }
    const double *a = scheme == H3D_EULER ? RK_A1 : scheme == H3D_RK3 ? RK_A3 : scheme == H3D_RK5 ? RK_A5 : RK_A14;
    const double *b = scheme == H3D_EULER ? RK_B1 : scheme == H3D_RK3 ? RK_B3 : scheme == H3D_RK5 ? RK_B5 : RK_B14;
    const double *c = scheme == H3D_EULER ? RK_C1 : scheme == H3D_RK3 ? RK_C3 : scheme == H3D_RK5 ? RK_C5 : RK_C14;
    const double tk = t + b[k] * dt;
    computeTimeDerivative(o, tk);
    if (scheme == H3D_EULER) {   // Q = Q + deltaT QDot (:1262-1271)
#pragma omp parallel for schedule(static)
        for (size_t q = 0; q < o.Q.size(); ++q) o.Q[q] = o.Q[q] + dt * o.QDot[q];
        return tk;
    }
    const double ak = a[k], cdt = c[k] * dt;
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < o.Q.size(); ++q) {
        o.G[q] = ak * o.G[q] + o.QDot[q];
        o.Q[q] = o.Q[q] + cdt * o.G[q];
    }
    return tk;
}
}  // namespace

int orc_enable_limiter(void* p, int enabled, double minimum) {
    Oracle& o = *(Oracle*)p;
    if (enabled && o.volume.empty()) { o.err = "the limiter needs the element volumes"; return 1; }
    o.limited = enabled != 0; if (minimum > 0.0) o.limiterMin = minimum;
    return 0;
}

int orc_rk_step(void* p, int scheme, double t, double dt, int ctd_after_step) {
    Oracle& o = *(Oracle*)p;
    const int ns = rkStages(scheme);
    if (!ns) { o.err = "unknown RK scheme"; return 1; }
    double tk = t;
    for (int k = 0; k < ns; ++k) tk = rkStage(o, scheme, k, t, dt);
    if (ctd_after_step) computeTimeDerivative(o, (scheme == H3D_RK3 || scheme == H3D_SSPRK33 || scheme == H3D_SSPRK43) ? t + dt : tk);
    return 0;
}

int orc_rk_stage(void* p, int scheme, int stage, double t, double dt) {
    Oracle& o = *(Oracle*)p;
    const int ns = rkStages(scheme);
    if (!ns) { o.err = "unknown RK scheme"; return 1; }
    if (stage < 0 || stage >= ns) { o.err = "Runge-Kutta stage out of range"; return 1; }
    rkStage(o, scheme, stage, t, dt);
    return 0;
}

// ComputeMaxResiduals (libs/discretization/DGSEMClass.f90:770-856)
int orc_max_residuals(void* p, double out[5]) {
    Oracle& o = *(Oracle*)p;
    double R[5] = {0, 0, 0, 0, 0};
    for (size_t g = 0; g < o.QDot.size() / 5; ++g) for (int q = 0; q < 5; ++q) R[q] = std::fmax(R[q], std::fabs(o.QDot[5 * g + q]));
    for (int q = 0; q < 5; ++q) out[q] = R[q];
    return 0;
}

// MaxTimeStep (DGSEMClass.f90:870-1034) with ComputeEigenvaluesForState (Physics_NS.f90:927-962)
int orc_max_timestep(void* p, double cfl, double dcfl, double* dt_conv, double* dt_visc) {
    Oracle& o = *(Oracle*)p;
    double TimeStep_Conv = std::numeric_limits<double>::max(), TimeStep_Visc = std::numeric_limits<double>::max();
    double dcsi = o.N != 0 ? 1.0 / std::fabs(o.x[1] - o.x[0]) : 0.0, deta = dcsi, dzet = dcsi;
    double dcsi2 = dcsi * dcsi, deta2 = deta * deta, dzet2 = dzet * dzet;
    const size_t nn = o.mixed ? PD(o).eOff[o.nElem] : (size_t)o.n3() * o.nElem;
    int eCur = -1;
    for (size_t g = 0; g < nn; ++g) {
        if (o.mixed) {   // the spacings of the element's own nodal storages (:915-939)
            const PData& P = PD(o);
            while (g >= P.eOff[eCur + 1]) {
                ++eCur;
                auto spacing = [&](int N) { const Basis1& b = P.sp.at(N); return N != 0 ? 1.0 / std::fabs(b.x[1] - b.x[0]) : 0.0; };
                dcsi = spacing(P.Nxyz[3 * eCur]); deta = spacing(P.Nxyz[3 * eCur + 1]); dzet = spacing(P.Nxyz[3 * eCur + 2]);
                dcsi2 = dcsi * dcsi; deta2 = deta * deta; dzet2 = dzet * dzet;
            }
        }
        const double* Q = &o.Q[5 * g];
        double u = std::fabs(Q[1] / Q[0]), v = std::fabs(Q[2] / Q[0]), w = std::fabs(Q[3] / Q[0]);
        double pr = Pressure(o, Q);
        double a = std::sqrt(o.ph.gamma * pr / Q[0]);
        double ev[3] = {u + a, v + a, w + a};
        double jac = o.jac[g];
        const double* jx = &o.JaXi[3 * g]; const double* je = &o.JaEta[3 * g]; const double* jz = &o.JaZeta[3 * g];
        double lamcsi_a = std::fabs(jx[0] * ev[0] + jx[1] * ev[1] + jx[2] * ev[2]) * dcsi;
        double lameta_a = std::fabs(je[0] * ev[0] + je[1] * ev[1] + je[2] * ev[2]) * deta;
        double lamzet_a = std::fabs(jz[0] * ev[0] + jz[1] * ev[1] + jz[2] * ev[2]) * dzet;
        TimeStep_Conv = std::fmin(TimeStep_Conv, cfl * std::fabs(jac) / (lamcsi_a + lameta_a + lamzet_a));
        if (o.ph.flowIsNavierStokes) {
            double T = Temperature(o, Q);
            double mu = SutherlandsLaw(o, T);
            double lamcsi_v = mu * dcsi2 * std::fabs(jx[0] + jx[1] + jx[2]);
            double lameta_v = mu * deta2 * std::fabs(je[0] + je[1] + je[2]);
            double lamzet_v = mu * dzet2 * std::fabs(jz[0] + jz[1] + jz[2]);
            TimeStep_Visc = std::fmin(TimeStep_Visc, dcfl * std::fabs(jac) / (lamcsi_v + lameta_v + lamzet_v));
        }
    }
    *dt_conv = TimeStep_Conv; *dt_visc = TimeStep_Visc;
    return 0;
}

// ScalarVolumeIntegral (libs/monitors/VolumeIntegrals.f90:76-120, 167-286)
// ScalarSurfaceIntegral / VectorSurfaceIntegral (libs/monitors/SurfaceIntegrals.f90:40-445) with getStressTensor
// (Physics_NS.f90:822-886); the state (and gradients) are prolonged anew as the reference does (:57-77)
int orc_surface_integral(void* p, int zone, int kind, double* out) {
    Oracle& o = *(Oracle*)p; Idx ix{o.n};
    if (kind < H3D_SURF_SURFACE || kind > H3D_SURF_VISCOUS_FORCE) { o.err = "unknown surface integral"; return 1; }
    const bool viscous = kind == H3D_SURF_TOTAL_FORCE || kind == H3D_SURF_VISCOUS_FORCE;
    if (viscous && !o.ph.computeGradients) { o.err = "surface integral needs gradients"; return 1; }
    if (o.mixed) {
        prolongToFacesP(o, 5, o.Q, o.fQ);
        if (o.ph.computeGradients) { prolongToFacesP(o, 5, o.Ux, o.fUx); prolongToFacesP(o, 5, o.Uy, o.fUy); prolongToFacesP(o, 5, o.Uz, o.fUz); }
    } else {
        prolongToFaces(o, 5, o.Q, o.fQ);
        if (o.ph.computeGradients) { prolongToFaces(o, 5, o.Ux, o.fUx); prolongToFaces(o, 5, o.Uy, o.fUy); prolongToFaces(o, 5, o.Uz, o.fUz); }
    }
    double val[3] = {0, 0, 0};
    for (int f = 0; f < o.nFace; ++f) {
        if (o.faceType[f] != H3D_FACE_BOUNDARY || o.faceZone[f] != zone) continue;
        double fv[3] = {0, 0, 0};
        // a p-nonconforming mesh integrates with the nodal storages of the face orders (SurfaceIntegrals.f90:100-104)
        const int n1 = o.mixed ? PD(o).fo[6 * f] + 1 : o.n, n2 = o.mixed ? PD(o).fo[6 * f + 1] + 1 : o.n;
        const double* w1 = o.mixed ? PD(o).sp.at(n1 - 1).w.data() : o.w.data(); const double* w2 = o.mixed ? PD(o).sp.at(n2 - 1).w.data() : o.w.data();
        for (int j = 0; j < n2; ++j) for (int i = 0; i < n1; ++i) {
            const size_t gL = o.mixed ? PD(o).fnode(f, 0, i, j) : ix.fnode(f, 0, i, j), gg = o.mixed ? PD(o).gnode(f, i, j) : ix.gnode(f, i, j);
            const double* Q = &o.fQ[gL * 5];
            const double* nh = &o.fNormal[3 * gg];
            const double Jf = o.fJac[gg];
            switch (kind) {
                case H3D_SURF_SURFACE: fv[0] = fv[0] + w1[i] * w2[j] * Jf; break;
                case H3D_SURF_MASS_FLOW: fv[0] = fv[0] + (Q[IRHOU] * nh[0] + Q[IRHOV] * nh[1] + Q[IRHOW] * nh[2]) * w1[i] * w2[j] * Jf; break;
                case H3D_SURF_FLOW_RATE: fv[0] = fv[0] + (1.0 / Q[IRHO]) * (Q[IRHOU] * nh[0] + Q[IRHOV] * nh[1] + Q[IRHOW] * nh[2]) * w1[i] * w2[j] * Jf; break;
                case H3D_SURF_PRESSURE: { double pr = Pressure(o, Q); fv[0] = fv[0] + pr * w1[i] * w2[j] * Jf; } break;
                case H3D_SURF_VEC_SURFACE: for (int d = 0; d < 3; ++d) fv[d] = fv[d] + w1[i] * w2[j] * Jf * nh[d]; break;
                case H3D_SURF_PRESSURE_FORCE: { double pr = Pressure(o, Q); for (int d = 0; d < 3; ++d) fv[d] = fv[d] + (pr * nh[d]) * Jf * w1[i] * w2[j]; } break;
                default: {
                    double U_x[3], U_y[3], U_z[3], tau[3][3], mu, kappa;
                    const double* gx = &o.fUx[gL * 5]; const double* gy = &o.fUy[gL * 5]; const double* gz = &o.fUz[gL * 5];
                    if (o.ph.gradientVariables == H3D_GRADVARS_ENTROPY) {   // getStressTensor's own entropy branch (Physics_NS.f90:858-866)
                        double invRho = 1.0 / Q[IRHO];
                        double p_div_rho = o.ph.gammaMinus1 * invRho * (Q[IRHOE] - 0.5 * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) * invRho);
                        for (int c = 0; c < 3; ++c) {
                            double u = Q[IRHOU + c] * invRho;
                            U_x[c] = p_div_rho * (gx[IRHOU + c] + u * gx[IRHOE]); U_y[c] = p_div_rho * (gy[IRHOU + c] + u * gy[IRHOE]); U_z[c] = p_div_rho * (gz[IRHOU + c] + u * gz[IRHOE]);
                        }
                    } else {
                        getVelocityGradients(o, Q, gx, gy, gz, U_x, U_y, U_z);   // STATE and ENERGY branches (:841-856) coincide with the pointer's
                    }
                    get_laminar_mu_kappa(o, Q, mu, kappa);      // mu = mu0 * SutherlandsLaw(T)
                    double divV = U_x[IX] + U_y[IY] + U_z[IZ];
                    tau[IX][IX] = mu * (2.0 * U_x[IX] - 2.0 / 3.0 * divV);
                    tau[IY][IX] = mu * (U_x[IY] + U_y[IX]);
                    tau[IZ][IX] = mu * (U_x[IZ] + U_z[IX]);
                    tau[IX][IY] = tau[IY][IX];
                    tau[IY][IY] = mu * (2.0 * U_y[IY] - 2.0 / 3.0 * divV);
                    tau[IZ][IY] = mu * (U_y[IZ] + U_z[IY]);
                    tau[IX][IZ] = tau[IZ][IX];
                    tau[IY][IZ] = tau[IZ][IY];
                    tau[IZ][IZ] = mu * (2.0 * U_z[IZ] - 2.0 / 3.0 * divV);
                    double pr = Pressure(o, Q);
                    for (int d = 0; d < 3; ++d) {
                        double tn = tau[d][0] * nh[0] + tau[d][1] * nh[1] + tau[d][2] * nh[2];   // matmul(tau, n)
                        if (kind == H3D_SURF_TOTAL_FORCE) fv[d] = fv[d] + (pr * nh[d] - tn) * Jf * w1[i] * w2[j];
                        else fv[d] = fv[d] - tn * Jf * w1[i] * w2[j];
                    }
                }
            }
        }
        for (int d = 0; d < 3; ++d) val[d] = val[d] + fv[d];
    }
    for (int d = 0; d < 3; ++d) out[d] = val[d];
    return 0;
}

// Probe_Update (libs/monitors/Probe.f90:330-420)
int orc_probe(void* p, int nProbes, const int* elem, const int* variable, const double* lxi, const double* leta, const double* lzeta, double* values) {
    Oracle& o = *(Oracle*)p; const int n = o.n; Idx ix{n};
    const double gamma = o.ph.gamma;
    // p-nonconforming meshes: the Lagrange vectors of probe pr are lxi[pr*ld + i] with ld = the largest number of nodes per direction
    int ld = n;
    if (o.mixed) { ld = 0; for (int q : PD(o).Nxyz) ld = std::max(ld, q + 1); }
    for (int pr = 0; pr < nProbes; ++pr) {
        if (elem[pr] < 0 || elem[pr] >= o.nElem) { o.err = "probe element out of range"; return 1; }
        double value = 0.0;
        const int nx = o.mixed ? PD(o).Nxyz[3 * elem[pr]] + 1 : n, ny = o.mixed ? PD(o).Nxyz[3 * elem[pr] + 1] + 1 : n, nz = o.mixed ? PD(o).Nxyz[3 * elem[pr] + 2] + 1 : n;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            const double* Q = &o.Q[5 * (o.mixed ? PD(o).eOff[elem[pr]] + (size_t)(k * ny + j) * nx + i : ix.node(elem[pr], i, j, k))];
            double var;
            switch (variable[pr]) {
                case H3D_PROBE_PRESSURE: var = Pressure(o, Q); break;
                case H3D_PROBE_VELOCITY: var = std::sqrt(POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) / Q[IRHO]; break;
                case H3D_PROBE_U: var = Q[IRHOU] / Q[IRHO]; break;
                case H3D_PROBE_V: var = Q[IRHOV] / Q[IRHO]; break;
                case H3D_PROBE_W: var = Q[IRHOW] / Q[IRHO]; break;
                case H3D_PROBE_MACH: {   // as written in the reference (:359-362), including its operator precedence
                    var = POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW]) / POW2(Q[IRHO]);
                    var = std::sqrt(var / (gamma * (gamma - 1.0) * (Q[IRHOE] / Q[IRHO] - 0.5 * var)));
                } break;
                case H3D_PROBE_K: var = 0.5 * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) / Q[IRHO]; break;
                default: o.err = "unknown probe variable"; return 1;
            }
            value = value + var * lxi[pr * ld + i] * leta[pr * ld + j] * lzeta[pr * ld + k];
        }
        values[pr] = value;
    }
    return 0;
}

// StatisticsMonitor_UpdateValues (libs/monitors/StatisticsMonitor.f90:279-540), velocities, Reynolds stresses, state [, gradients]
int orc_snapshot_begin(void* p) { Oracle& o = *(Oracle*)p; o.snapshot = o.Q; return 0; }
int orc_snapshot_end(void* p, double* Q) {
    Oracle& o = *(Oracle*)p;
    if (o.snapshot.empty()) { o.err = "no snapshot in flight"; return 1; }
    if (Q) std::memcpy(Q, o.snapshot.data(), o.snapshot.size() * sizeof(double));
    o.snapshot.clear();
    return 0;
}
int orc_statistics_update(void* p, int reset) {
    Oracle& o = *(Oracle*)p;
    const int nv = o.ph.computeGradients ? 29 : 14;
    const size_t nn = o.mixed ? PD(o).eOff[o.nElem] : (size_t)o.nElem * o.n3();
    if (reset || o.stats.size() != nn * nv) { o.stats.assign(nn * nv, 0.0); o.statSamples = 0; }
    const double inv_nsamples_plus_1 = 1.0 / (o.statSamples + 1);
    const double ratio = o.statSamples * inv_nsamples_plus_1;
    for (size_t g = 0; g < nn; ++g) {
        double* data = &o.stats[g * nv]; const double* Q = &o.Q[5 * g];
        const double rfactor1 = inv_nsamples_plus_1 / Q[IRHO], rfactor2 = inv_nsamples_plus_1 / POW2(Q[IRHO]);
        data[0] = data[0] * ratio + Q[IRHOU] * rfactor1;
        data[1] = data[1] * ratio + Q[IRHOV] * rfactor1;
        data[2] = data[2] * ratio + Q[IRHOW] * rfactor1;
        data[3] = data[3] * ratio + POW2(Q[IRHOU]) * rfactor2;
        data[4] = data[4] * ratio + POW2(Q[IRHOV]) * rfactor2;
        data[5] = data[5] * ratio + POW2(Q[IRHOW]) * rfactor2;
        data[6] = data[6] * ratio + Q[IRHOU] * Q[IRHOV] * rfactor2;
        data[7] = data[7] * ratio + Q[IRHOU] * Q[IRHOW] * rfactor2;
        data[8] = data[8] * ratio + Q[IRHOV] * Q[IRHOW] * rfactor2;
        for (int q = 0; q < 5; ++q) data[9 + q] = data[9 + q] * ratio + Q[q] * inv_nsamples_plus_1;
        if (nv == 29) for (int q = 0; q < 5; ++q) {
            data[14 + q] = data[14 + q] * ratio + o.Ux[5 * g + q] * inv_nsamples_plus_1;
            data[19 + q] = data[19 + q] * ratio + o.Uy[5 * g + q] * inv_nsamples_plus_1;
            data[24 + q] = data[24 + q] * ratio + o.Uz[5 * g + q] * inv_nsamples_plus_1;
        }
    }
    ++o.statSamples;
    return 0;
}
int orc_statistics_download(void* p, double* data, int* nVars, int* nSamples) {
    Oracle& o = *(Oracle*)p;
    if (o.stats.empty()) { o.err = "no statistics have been accumulated"; return 1; }
    *nVars = (int)(o.stats.size() / (o.mixed ? PD(o).eOff[o.nElem] : (size_t)o.nElem * o.n3())); *nSamples = o.statSamples;
    if (data) std::memcpy(data, o.stats.data(), o.stats.size() * sizeof(double));
    return 0;
}

int orc_volume_integral(void* p, int kind, double* out) {
    Oracle& o = *(Oracle*)p; const int n = o.n; Idx ix{n};
    double val = 0.0;
    for (int e = 0; e < o.nElem; ++e) {
        double loc = 0.0;
        // p-nonconforming meshes: the element's own nodal storages (VolumeIntegrals.f90:190-196)
        const int nx = o.mixed ? PD(o).Nxyz[3 * e] + 1 : n, ny = o.mixed ? PD(o).Nxyz[3 * e + 1] + 1 : n, nz = o.mixed ? PD(o).Nxyz[3 * e + 2] + 1 : n;
        const double* wx = o.mixed ? PD(o).sp.at(nx - 1).w.data() : o.w.data(); const double* wy = o.mixed ? PD(o).sp.at(ny - 1).w.data() : o.w.data();
        const double* wz = o.mixed ? PD(o).sp.at(nz - 1).w.data() : o.w.data();
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            size_t g = o.mixed ? PD(o).eOff[e] + (size_t)(k * ny + j) * nx + i : ix.node(e, i, j, k);
            const double* Q = &o.Q[5 * g]; const double* QD = &o.QDot[5 * g];
            double wJ = wx[i] * wy[j] * wz[k] * o.jac[g];
            switch (kind) {
                case H3D_INT_VOLUME: loc = loc + wJ; break;
                case H3D_INT_KINETIC_ENERGY: {
                    double KinEn = POW2(Q[IRHOU]); KinEn = KinEn + POW2(Q[IRHOV]); KinEn = KinEn + POW2(Q[IRHOW]);
                    KinEn = 0.5 * KinEn / Q[IRHO];
                    loc = loc + wJ * KinEn;
                } break;
                case H3D_INT_KINETIC_ENERGY_RATE: {
                    double uvw = Q[IRHOU] / Q[IRHO];
                    double KinEn = uvw * QD[IRHOU] - 0.5 * POW2(uvw) * QD[IRHO];
                    uvw = Q[IRHOV] / Q[IRHO];
                    KinEn = KinEn + uvw * QD[IRHOV] - 0.5 * POW2(uvw) * QD[IRHO];
                    uvw = Q[IRHOW] / Q[IRHO];
                    KinEn = KinEn + uvw * QD[IRHOW] - 0.5 * POW2(uvw) * QD[IRHO];
                    loc = loc + wJ * KinEn;
                } break;
                case H3D_INT_ENSTROPHY: {
                    double U_x[3], U_y[3], U_z[3];
                    getVelocityGradients(o, Q, &o.Ux[5 * g], &o.Uy[5 * g], &o.Uz[5 * g], U_x, U_y, U_z);
                    double KinEn = POW2(U_y[IZ] - U_z[IY]) + POW2(U_z[IX] - U_x[IZ]) + POW2(U_x[IY] - U_y[IX]);
                    loc = loc + wJ * KinEn;
                } break;
                case H3D_INT_KINETIC_ENERGY_BALANCE: {
                    // kinetic energy rate + viscous work - pressure work + de-aliasing correction (:220-265)
                    auto nodeAt = [&](int a, int b, int c) { return o.mixed ? PD(o).eOff[e] + (size_t)(c * ny + b) * nx + a : ix.node(e, a, b, c); };
                    auto pressureAt = [&](int a, int b, int c) { return Pressure(o, &o.Q[5 * nodeAt(a, b, c)]); };
                    const double* Dx = o.mixed ? PD(o).sp.at(nx - 1).D.data() : o.D.data(); const double* Dy = o.mixed ? PD(o).sp.at(ny - 1).D.data() : o.D.data();
                    const double* Dz = o.mixed ? PD(o).sp.at(nz - 1).D.data() : o.D.data();
                    double grad_Mp[3] = {0, 0, 0}, M_grad_p[3] = {0, 0, 0};     // GetPressureLocalGradient (:724-764)
                    for (int l = 0; l < nx; ++l) { double pl = pressureAt(l, j, k); size_t gl = nodeAt(l, j, k);
                        for (int d = 0; d < 3; ++d) { grad_Mp[d] = grad_Mp[d] + pl * o.JaXi[3 * gl + d] * Dx[i * nx + l]; M_grad_p[d] = M_grad_p[d] + pl * o.JaXi[3 * g + d] * Dx[i * nx + l]; } }
                    for (int l = 0; l < ny; ++l) { double pl = pressureAt(i, l, k); size_t gl = nodeAt(i, l, k);
                        for (int d = 0; d < 3; ++d) { grad_Mp[d] = grad_Mp[d] + pl * o.JaEta[3 * gl + d] * Dy[j * ny + l]; M_grad_p[d] = M_grad_p[d] + pl * o.JaEta[3 * g + d] * Dy[j * ny + l]; } }
                    for (int l = 0; l < nz; ++l) { double pl = pressureAt(i, j, l); size_t gl = nodeAt(i, j, l);
                        for (int d = 0; d < 3; ++d) { grad_Mp[d] = grad_Mp[d] + pl * o.JaZeta[3 * gl + d] * Dz[k * nz + l]; M_grad_p[d] = M_grad_p[d] + pl * o.JaZeta[3 * g + d] * Dz[k * nz + l]; } }
                    double inv_rho = 1.0 / Q[IRHO];
                    double uvw = Q[IRHOU] * inv_rho;
                    double KinEn = uvw * QD[IRHOU] - 0.5 * POW2(uvw) * QD[IRHO];
                    uvw = Q[IRHOV] * inv_rho; KinEn = KinEn + uvw * QD[IRHOV] - 0.5 * POW2(uvw) * QD[IRHO];
                    uvw = Q[IRHOW] * inv_rho; KinEn = KinEn + uvw * QD[IRHOW] - 0.5 * POW2(uvw) * QD[IRHO];
                    double p3 = o.ph.gammaMinus1 * (Q[IRHOE] - 0.5 * (POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) * inv_rho);
                    double corr = 0.5 * (Q[IRHOU] * (M_grad_p[IX] - grad_Mp[IX]) + Q[IRHOV] * (M_grad_p[IY] - grad_Mp[IY]) + Q[IRHOW] * (M_grad_p[IZ] - grad_Mp[IZ])) * inv_rho;
                    double F[NCONS][NDIM];
                    const double* gx = &o.Ux[5 * g]; const double* gy = &o.Uy[5 * g]; const double* gz = &o.Uz[5 * g];
                    ViscousFlux_ENERGY(o, Q, gx, gy, gz, o.mu[2 * g], 0.0, o.mu[2 * g + 1], F);
                    double work = 0.0;
                    for (int q = IRHOU; q <= IRHOW; ++q) work = work + (F[q][IX] * gx[q] + F[q][IY] * gy[q] + F[q][IZ] * gz[q]);
                    loc = loc + wx[i] * wy[j] * wz[k] * (o.jac[g] * (KinEn + work - p3 * (gx[IRHOU] + gy[IRHOV] + gz[IRHOW])) + corr);
                } break;
                case H3D_INT_VELOCITY:
                    loc = loc + wx[i] * wy[j] * wz[k] * std::sqrt(POW2(Q[IRHOU]) + POW2(Q[IRHOV]) + POW2(Q[IRHOW])) / Q[IRHO] * o.jac[g];
                    break;
                case H3D_INT_ENTROPY: {
                    double pr = Pressure(o, Q);
                    double sp = std::log(pr) - o.ph.gamma * std::log(Q[IRHO]);
                    loc = loc + wJ * sp;
                } break;
                case H3D_INT_MATH_ENTROPY: {
                    double pr = Pressure(o, Q);
                    double sp = std::log(pr) - o.ph.gamma * std::log(Q[IRHO]);
                    double ms = -Q[IRHO] * sp / o.ph.gammaMinus1;
                    loc = loc + wJ * ms;
                } break;
                case H3D_INT_INTERNAL_ENERGY: loc = loc + wJ * Q[IRHOE]; break;
                case H3D_INT_ENTROPY_RATE: case H3D_INT_ENTROPY_BALANCE: {
                    // NSGradientVariables_ENTROPY whatever the gradient variables of the run (:326, :343, :353, :363)
                    double EV[5];
                    GetGradientsAs(o.ph, H3D_GRADVARS_ENTROPY, Q, EV);
                    double dot = 0.0;
                    for (int q = 0; q < 5; ++q) dot = dot + QD[q] * EV[q];
                    if (kind == H3D_INT_ENTROPY_BALANCE) {
                        double F[NCONS][NDIM], work = 0.0;
                        const double* gx = &o.Ux[5 * g]; const double* gy = &o.Uy[5 * g]; const double* gz = &o.Uz[5 * g];
                        ViscousFlux(o, Q, gx, gy, gz, o.mu[2 * g], 0.0, o.mu[2 * g + 1], F);
                        for (int q = 0; q < 5; ++q) work = work + (F[q][IX] * gx[q] + F[q][IY] * gy[q] + F[q][IZ] * gz[q]);
                        dot = dot + work;
                    }
                    loc = loc + wJ * dot;
                } break;
                default: o.err = "unknown volume integral"; return 1;
            }
        }
        val = val + loc;
    }
    *out = val;
    return 0;
}

// checkForNan (ExplicitMethods.f90:1856-1905)
int orc_has_nan(void* p, int* flag) {
    Oracle& o = *(Oracle*)p; int f = 0;
    for (double q : o.Q) if (std::isnan(q)) { f = 1; break; }
    *flag = f;
    return 0;
}

// ---- independent restatement of the 1-D operators for the K6 pin (NodalStorageClass.f90:201-275) -------
// LegendreAlgorithms.f90:137-212 (Gauss), :275-358 (Lobatto); InterpolationAndDerivatives.f90:230-255, 799-828, 109-159
static void orc_legendre(int N, double x, double& L, double& dL) {
    if (N == 0) { L = 1; dL = 0; return; }
    if (N == 1) { L = x; dL = 1; return; }
    double Lm2 = 1, dLm2 = 0, Lm1 = x, dLm1 = 1; L = 0; dL = 0;
    for (int k = 2; k <= N; ++k) { L = ((2 * k - 1) * x * Lm1 - (k - 1) * Lm2) / k; dL = dLm2 + (2 * k - 1) * Lm1; Lm2 = Lm1; Lm1 = L; dLm2 = dLm1; dLm1 = dL; }
}
int orc_nodal(int N, int nodeType, double* x, double* w, double* D, double* hatD, double* sharpD, double* v, double* b) {
    const int n = N + 1; const double PI = 3.141592653589793238462643, tol = 4.0 * std::numeric_limits<double>::epsilon();
    if (nodeType == H3D_GAUSS) {
        if (N == 0) { x[0] = 0; w[0] = 2; }
        else if (N == 1) { x[0] = -std::sqrt(1.0 / 3.0); w[0] = 1; x[1] = -x[0]; w[1] = 1; }
        else for (int j = 0; j < (N + 1) / 2; ++j) {
            double xj = -std::cos((2 * j + 1) * PI / (2 * N + 2)), L, dL;
            for (int k = 0; k <= 10; ++k) { orc_legendre(N + 1, xj, L, dL); double d = -L / dL; xj += d; if (std::fabs(d) <= tol * std::fabs(xj)) break; }
            orc_legendre(N + 1, xj, L, dL);
            x[j] = xj; w[j] = 2.0 / ((1.0 - xj * xj) * dL * dL); x[N - j] = -xj; w[N - j] = w[j];
        }
        if (N % 2 == 0 && N > 0) { double L, dL; orc_legendre(N + 1, 0.0, L, dL); x[N / 2] = 0; w[N / 2] = 2.0 / (dL * dL); }
    } else {
        if (N == 1) { x[0] = -1; w[0] = 1; x[1] = 1; w[1] = 1; }
        else {
            x[0] = -1; w[0] = 2.0 / (N * (N + 1)); x[N] = 1; w[N] = w[0];
            auto qAndL = [&](double xx, double& Q, double& dQ, double& LN) {
                double Lm2 = 1, dLm2 = 0, Lm1 = xx, dLm1 = 1, Lk = 0, dLk = 0;
                for (int k = 2; k <= N; ++k) { Lk = ((2 * k - 1) * xx * Lm1 - (k - 1) * Lm2) / k; dLk = dLm2 + (2 * k - 1) * Lm1; Lm2 = Lm1; Lm1 = Lk; dLm2 = dLm1; dLm1 = dLk; }
                int k = N + 1; Lk = ((2 * k - 1) * xx * Lm1 - (k - 1) * Lm2) / k; dLk = dLm2 + (2 * k - 1) * Lm1;
                Q = Lk - Lm2; dQ = dLk - dLm2; LN = Lm1;
            };
            for (int j = 1; j < (N + 1) / 2; ++j) {
                double xj = -std::cos((j + 0.25) * PI / N - 3.0 / (8 * N * PI * (j + 0.25))), Q, dQ, LN;
                for (int k = 0; k <= 10; ++k) { qAndL(xj, Q, dQ, LN); double d = -Q / dQ; xj += d; if (std::fabs(d) <= tol * std::fabs(xj)) break; }
                qAndL(xj, Q, dQ, LN);
                x[j] = xj; w[j] = 2.0 / (N * (N + 1) * LN * LN); x[N - j] = -xj; w[N - j] = w[j];
            }
            if (N % 2 == 0) { double L, dL; orc_legendre(N, 0.0, L, dL); x[N / 2] = 0; w[N / 2] = 2.0 / (N * (N + 1) * L * L); }
        }
    }
    std::vector<double> wb(n, 1.0);
    for (int j = 1; j <= N; ++j) for (int k = 0; k < j; ++k) { wb[k] *= (x[k] - x[j]); wb[j] *= (x[j] - x[k]); }
    for (int j = 0; j <= N; ++j) wb[j] = 1.0 / wb[j];
    for (int i = 0; i <= N; ++i) { D[i * n + i] = 0; for (int j = 0; j <= N; ++j) if (j != i) { D[i * n + j] = wb[j] / (wb[i] * (x[i] - x[j])); D[i * n + i] -= D[i * n + j]; } }
    for (int j = 0; j <= N; ++j) for (int i = 0; i <= N; ++i) hatD[i * n + j] = D[j * n + i] * w[j] / w[i];
    for (int q = 0; q < n * n; ++q) sharpD[q] = 0.0;
    if (nodeType == H3D_GAUSSLOBATTO && N != 0) { for (int q = 0; q < n * n; ++q) sharpD[q] = 2.0 * D[q]; sharpD[0] = 2.0 * D[0] + 1.0 / w[0]; sharpD[N * n + N] = 2.0 * D[N * n + N] - 1.0 / w[N]; }
    for (int s = 0; s < 2; ++s) {
        double xe = s ? 1.0 : -1.0; bool match = false;
        for (int j = 0; j <= N; ++j) { v[s * n + j] = 0; double a = xe, bb = x[j]; double tl = 2 * std::numeric_limits<double>::epsilon();
            bool eq = (a == 0.0 || bb == 0.0) ? std::fabs(a - bb) <= tl : std::fabs(bb - a) <= tl * std::fmax(std::fabs(a), std::fabs(bb));
            if (eq) { v[s * n + j] = 1; match = true; } }
        if (!match) { double d = 0; for (int j = 0; j <= N; ++j) { double t = wb[j] / (xe - x[j]); v[s * n + j] = t; d += t; } for (int j = 0; j <= N; ++j) v[s * n + j] /= d; }
        for (int j = 0; j <= N; ++j) b[s * n + j] = v[s * n + j] / w[j];
    }
    return 0;
}

}  // extern "C"
