// p-nonconforming meshes on the device: every element has its own polynomial orders (Nx, Ny, Nz), every face the
// orders of its two sides and its own, Nf = max per direction.  The element traces are interpolated to the face order, the
// interface fluxes are computed there and projected back to each element (mortars).
//
//   Face_LinkWithElements                  FaceClass.f90:187-282     (orders, projection types: host, MixedSolver::setMesh)
//   HexElement_ProlongSolutionToFaces      HexElementClass.f90:233-372   MxTrace  (element -> its own face order)
//   Face_AdaptSolutionToFace / ...Gradients FaceClass.f90:284-512    MxAdapt  (rotation + Tset interpolation to the face order)
//   HexElement_ComputeLocalGradient        HexElementClass.f90:427-531   MxLocalGrad
//   BR1_ComputeElementInterfaceAverage / BR1_ComputeBoundaryFlux  EllipticBR1.f90:571-736   MxGradFace
//   Face_ProjectFluxToElements / Face_ProjectGradientFluxToElements  FaceClass.f90:596-696, 865-961   MxProject
//   BR1_GradientFaceLoop -> VectorWeakIntegrals_StdFace   EllipticBR1.f90:531-569, DGIntegrals.f90:365-443   MxLift
//   BaseClass_ComputeInnerFluxes / BR1_ComputeInnerFluxes  HyperbolicDiscretizationClass.f90:83-152, EllipticBR1.f90:740-814   MxFlux
//   computeElementInterfaceFlux / computeBoundaryFlux     SpatialDiscretization.f90:1710-2028   MxRiemann
//   ScalarWeakIntegrals_StdVolumeGreen + StdFace, /J, +S, RK update   DGIntegrals.f90:56-87, 214-273   MxVolume
//
// Design: the orders are run-time data, so these are plain gather kernels -- one thread per output node (element node, element
// trace node or face node), every sum in the reference's order, no shared memory and no synchronisation; the uniform-order
// path (h3d_kernels.cuh) stays the tuned one.  Layout: structure of arrays over the CONCATENATED nodes of all elements / faces
// (A[c][node]), with offset tables per element, element side and face.  Every kernel body is a functor over the thread index
// and the orchestration is a template over a backend (allocate / copy / launch): libh3dgpu.so instantiates it with the CUDA
// backend; tests/emu instantiates the SAME functors and orchestration with a host loop as the launcher, which is how this path
// is checked against the oracle where no GPU is present (test infrastructure, never shipped).
// Scope: StandardDG and SplitDG (any two-point flux), BR1 or interior penalty (or Euler), any Riemann solver / boundary condition / gradient variables / LES
// model of h3d_physics.cuh.
// Partitioned meshes: the traces of the MPI faces are exchanged at the face order (h3d_set_halo), scalars are all-reduced.  Reductions are computed per element (face) in the reference's node order and finished on the host in element
// order, so that they reproduce the oracle's sums bit for bit.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "h3d_phys_host.hpp"
#include "h3d_physics.cuh"

namespace h3d {

constexpr int MX_MAXN = 16;   // polynomial orders 0 .. 15

struct MixedDev {
    int nElem, nFace;
    long long nNodes, nFaceNodes, nTrace;
    // ---- tables
    const long long* eOff;       // [nElem + 1]      first node of every element
    const int* eN;               // [nElem][3]       nodes per direction (Nx+1, Ny+1, Nz+1)
    const int* nodeElem;         // [nNodes]         element of every node
    const long long* tOff;       // [6 nElem + 1]    first trace node of element side 6 e + lf (at the ELEMENT's face order)
    const int* traceOwner;       // [nTrace]         6 e + lf
    const long long* fOff;       // [nFace + 1]      first node of every face (at the FACE order)
    const int* faceNodeFace;     // [nFaceNodes]     face of every face node
    const int* fo;               // [nFace][6]       Nf(1:2), NfLeft(1:2), NfRight(1:2)
    const int* proj;             // [nFace][2]       projectionType of the two sides
    const int *elemFace, *elemFaceSide;     // [nElem][6]
    const int *faceElem, *faceElemSide;     // [nFace][2]
    const int *faceRot, *faceType, *faceZone;
    // ---- operators: per order N the block D[n n] hatD[n n] v[2 n] b[2 n] w[n] x[n] sharpD[n n] at ops + opBase[N] (row-major M[i n + l] = M(i,l))
    const double* ops; int opBase[MX_MAXN];
    // Tset(Norigin, Ndest) % T at tset + tBase[Norigin][Ndest], [(Ndest+1)][(Norigin+1)]
    const double* tset; int tBase[MX_MAXN][MX_MAXN];
    // ---- element fields [c][nNodes]
    double *Q, *G, *QDot, *Ux, *Uy, *Uz;    // 5 each
    double* Fc;                             // 15: contravariant flux (d*5 + q)
    const double* S;                        // 5 or nullptr
    const double *Ja, *J, *invJ;            // 9 (3 d + c), 1, 1
    const double *lesDelta, *fDelta;        // [nElem] (V / product(Nxyz+1))^(1/3), [nFace] sqrt(surface / product(Nf+1)) (SpatialDiscretization.f90:420, 1378)
    const double *dWall, *fDWall;           // [nNodes], [nFaceNodes] wall distances (LES wall model) or nullptr
    const double *volume;                   // [nElem] e % geom % volume (stage limiter) or nullptr
    const double *fPenalty;                 // [nFace] interior penalty: 1/2 sigma (max Nf + 1)(max Nf + 2) / h (EllipticIP.f90:678-687) or nullptr
    // ---- element-side fields at the element's face order [c][nTrace]
    double *tr;                             // 15: traces before the adaption to the face order
    double *fStarE, *unStarE;               // 5 / 15 (d*5 + q)
    // ---- face fields at the face order [c][nFaceNodes]
    double *fQ;                             // 10: side*5 + q
    double *fU;                             // 30: dir*10 + side*5 + q
    double *fFlux;                          // 15: interface flux (5) or gradient flux (d*5 + q)
    const double *fN, *fT1, *fT2, *fJ;      // 3, 3, 3, 1
    const int* bcType; const double* bcParams;
    // ---- reductions
    double* partial; long long partialStride;   // [8][partialStride], partialStride = max(nElem, nFace): column q of element / face e at q * stride + e
    // ---- MPI faces (h3d_set_halo): halo face k = haloFace[k], local side haloSide[k]; its nodes are the halo nodes
    //      [hOff[k], hOff[k+1]); neighbour b owns the halo nodes [nbrNodeOff[b], nbrNodeOff[b+1]) (faces in exchange order)
    int nHalo, nNbr; long long nHaloNodes;
    const int *haloFace, *haloSide, *haloNbr;   // [nHalo]
    const long long* hOff;                      // [nHalo + 1]
    const int* hNodeFace;                       // [nHaloNodes] halo face of every halo node
    const long long* nbrNodeOff;                // [nNbr + 1]
    double *sendBuf, *recvBuf;                  // [15 nHaloNodes]
};

struct MxRk { int mode; double a, cdt, b; int copyG; };   // as RkArgs (h3d_kernels.cuh): 0 residual only, 1 low storage, 2 SSP

// ---- index helpers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mxAxis0(int lf) { return (lf == 3 || lf == 5) ? 1 : 0; }            // axisMap(1, lf)
__device__ __forceinline__ int mxAxis1(int lf) { return (lf == 2 || lf == 4) ? 1 : 2; }            // axisMap(2, lf)
__device__ __forceinline__ int mxAxisN(int lf) { return (lf < 2) ? 1 : ((lf == 2 || lf == 4) ? 2 : 0); }
__device__ __forceinline__ int mxEnd(int lf) { return (lf == 1 || lf == 3 || lf == 4) ? 1 : 0; }
__device__ __forceinline__ const double* mxD(const MixedDev& m, int N) { return m.ops + m.opBase[N]; }
__device__ __forceinline__ const double* mxHatD(const MixedDev& m, int N) { return m.ops + m.opBase[N] + (N + 1) * (N + 1); }
__device__ __forceinline__ const double* mxV(const MixedDev& m, int N) { return m.ops + m.opBase[N] + 2 * (N + 1) * (N + 1); }
__device__ __forceinline__ const double* mxB(const MixedDev& m, int N) { return m.ops + m.opBase[N] + 2 * (N + 1) * (N + 1) + 2 * (N + 1); }
__device__ __forceinline__ const double* mxW(const MixedDev& m, int N) { return m.ops + m.opBase[N] + 2 * (N + 1) * (N + 1) + 4 * (N + 1); }
__device__ __forceinline__ const double* mxX(const MixedDev& m, int N) { return m.ops + m.opBase[N] + 2 * (N + 1) * (N + 1) + 5 * (N + 1); }
__device__ __forceinline__ const double* mxSharpD(const MixedDev& m, int N) { return m.ops + m.opBase[N] + 2 * (N + 1) * (N + 1) + 6 * (N + 1); }
__device__ __forceinline__ const double* mxT(const MixedDev& m, int No, int Nd) { return m.tset + m.tBase[No][Nd]; }
// MeshTypes.f90:70-108 and its inverse
__device__ __forceinline__ void mxLeft2Right(int i, int j, int Nx, int Ny, int rot, int& ii, int& jj) {
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
__device__ __forceinline__ void mxRight2Left(int ii, int jj, int Nx, int Ny, int rot, int& i, int& j) {
    switch (rot) {
        case 0: i = ii; j = jj; break;
        case 1: i = jj; j = Ny - ii; break;
        case 2: i = Nx - ii; j = Ny - jj; break;
        case 3: i = Nx - jj; j = ii; break;
        case 4: i = jj; j = ii; break;
        case 5: i = Nx - ii; j = jj; break;
        case 6: i = Nx - jj; j = Ny - ii; break;
        default: i = ii; j = Ny - jj; break;
    }
}
struct MxNode { int e, i, j, k, nx, ny, nz; long long base; };
__device__ __forceinline__ MxNode mxNode(const MixedDev& m, long long g) {
    MxNode t; t.e = m.nodeElem[g]; t.base = m.eOff[t.e];
    t.nx = m.eN[3 * t.e]; t.ny = m.eN[3 * t.e + 1]; t.nz = m.eN[3 * t.e + 2];
    const int l = (int)(g - t.base);
    t.i = l % t.nx; t.j = (l / t.nx) % t.ny; t.k = l / (t.nx * t.ny);
    return t;
}

// ---- HexElement_ProlongSolutionToFaces: trace of 5 fields on every element side, at the element's own face order ----------
struct MxTrace {
    MixedDev m; int nSets; const double* src[3]; double* dst;   // nSets fields of 5 (Q: 1; U_x, U_y, U_z: 3): src[s] [5][nNodes], dst [s*5 + c][nTrace]
    __device__ void operator()(long long t) const {
        const int owner = m.traceOwner[t], e = owner / 6, lf = owner % 6;
        const int nn[3] = {m.eN[3 * e], m.eN[3 * e + 1], m.eN[3 * e + 2]};
        const int a0 = mxAxis0(lf), a1 = mxAxis1(lf), an = mxAxisN(lf);
        const int local = (int)(t - m.tOff[owner]);
        int idx[3]; idx[a0] = local % nn[a0]; idx[a1] = local / nn[a0]; idx[an] = 0;
        const double* v = mxV(m, nn[an] - 1) + mxEnd(lf) * nn[an];
        const long long stride = an == 0 ? 1 : (an == 1 ? nn[0] : (long long)nn[0] * nn[1]);
        const long long g0 = m.eOff[e] + ((long long)idx[2] * nn[1] + idx[1]) * nn[0] + idx[0];
        for (int st = 0; st < nSets; ++st) {
            const double* f = src[st];
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            for (int l = 0; l < nn[an]; ++l) {
                const double vl = v[l];
                for (int c = 0; c < 5; ++c) acc[c] = acc[c] + f[(long long)c * m.nNodes + g0 + l * stride] * vl;
            }
            for (int c = 0; c < 5; ++c) dst[(long long)(st * 5 + c) * m.nTrace + t] = acc[c];
        }
    }
};

// ---- Face_AdaptSolutionToFace / ...GradientsToFace: the traces -> side `side` of the face storage at the face order -------------
struct MxAdapt {
    MixedDev m; int nSets; const double* src; double* dst;   // src [s*5 + c][nTrace]; dst [s*10 + side*5 + c][nFaceNodes]
    __device__ void operator()(long long t) const {
        const int side = (int)(t / m.nFaceNodes); const long long g = t % m.nFaceNodes;
        const int f = m.faceNodeFace[g];
        if (m.faceElem[2 * f + side] < 0) return;      // boundary face, or the remote side of an MPI face (filled by the exchange)
        const int* fo = m.fo + 6 * f;
        const int Nf1 = fo[0], Ns1 = fo[2 + 2 * side], Ns2 = fo[3 + 2 * side];
        const int local = (int)(g - m.fOff[f]), i = local % (Nf1 + 1), j = local / (Nf1 + 1);
        const int rot = side ? m.faceRot[f] : 0;
        const bool swap = rot == 1 || rot == 3 || rot == 4 || rot == 6;
        const int ne1 = (swap ? Ns2 : Ns1) + 1;
        const long long base = m.tOff[6 * m.faceElem[2 * f + side] + m.faceElemSide[2 * f + side]];
        auto at = [&](int a, int b) { int ii, jj; mxLeft2Right(a, b, Ns1, Ns2, rot, ii, jj); return base + (long long)jj * ne1 + ii; };
        const int pt = m.proj[2 * f + side];
        const double* T1 = (pt & 1) ? mxT(m, Ns1, Nf1) + i * (Ns1 + 1) : nullptr;
        const double* T2 = (pt & 2) ? mxT(m, Ns2, fo[1]) + j * (Ns2 + 1) : nullptr;
        for (int st = 0; st < nSets; ++st) {
            const double* sf = src + (long long)(st * 5) * m.nTrace;
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            if (pt == 0) {
                const long long s = at(i, j);
                for (int c = 0; c < 5; ++c) acc[c] = sf[(long long)c * m.nTrace + s];
            } else if (pt == 1) {
                for (int l = 0; l <= Ns1; ++l) { const long long s = at(l, j); for (int c = 0; c < 5; ++c) acc[c] = acc[c] + T1[l] * sf[(long long)c * m.nTrace + s]; }
            } else if (pt == 2) {
                for (int l = 0; l <= Ns2; ++l) { const long long s = at(i, l); for (int c = 0; c < 5; ++c) acc[c] = acc[c] + T2[l] * sf[(long long)c * m.nTrace + s]; }
            } else {
                for (int l = 0; l <= Ns2; ++l) for (int mm = 0; mm <= Ns1; ++mm) {
                    const long long s = at(mm, l); const double tt = T1[mm] * T2[l];
                    for (int c = 0; c < 5; ++c) acc[c] = acc[c] + tt * sf[(long long)c * m.nTrace + s];
                }
            }
            for (int c = 0; c < 5; ++c) dst[(long long)(st * 10 + side * 5 + c) * m.nFaceNodes + g] = acc[c];
        }
    }
};

// ---- HexElement_ComputeLocalGradient --------------------------------------------------------------------------------------
struct MxLocalGrad {
    MixedDev m; Phys ph;
    __device__ void operator()(long long g) const {
        const MxNode t = mxNode(m, g);
        const double* Dx = mxD(m, t.nx - 1) + t.i * t.nx; const double* Dy = mxD(m, t.ny - 1) + t.j * t.ny; const double* Dz = mxD(m, t.nz - 1) + t.k * t.nz;
        auto load = [&](long long gg, double* U) {
            double Q[5];
            for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * m.nNodes + gg];
            if (ph.gradVars == H3D_GRADVARS_STATE) { for (int q = 0; q < 5; ++q) U[q] = Q[q]; } else get_gradients(ph, Q, U);
        };
        double a[5] = {0, 0, 0, 0, 0}, b[5] = {0, 0, 0, 0, 0}, c[5] = {0, 0, 0, 0, 0}, U[5];
        for (int l = 0; l < t.nx; ++l) { load(t.base + ((long long)t.k * t.ny + t.j) * t.nx + l, U); for (int q = 0; q < 5; ++q) a[q] = a[q] + U[q] * Dx[l]; }
        for (int l = 0; l < t.ny; ++l) { load(t.base + ((long long)t.k * t.ny + l) * t.nx + t.i, U); for (int q = 0; q < 5; ++q) b[q] = b[q] + U[q] * Dy[l]; }
        for (int l = 0; l < t.nz; ++l) { load(t.base + ((long long)l * t.ny + t.j) * t.nx + t.i, U); for (int q = 0; q < 5; ++q) c[q] = c[q] + U[q] * Dz[l]; }
        double ja[9];
        for (int q = 0; q < 9; ++q) ja[q] = m.Ja[(long long)q * m.nNodes + g];
        const double iJ = m.invJ[g];
        for (int q = 0; q < 5; ++q) {
            m.Ux[(long long)q * m.nNodes + g] = (a[q] * ja[0] + b[q] * ja[3] + c[q] * ja[6]) * iJ;
            m.Uy[(long long)q * m.nNodes + g] = (a[q] * ja[1] + b[q] * ja[4] + c[q] * ja[7]) * iJ;
            m.Uz[(long long)q * m.nNodes + g] = (a[q] * ja[2] + b[q] * ja[5] + c[q] * ja[8]) * iJ;
        }
    }
};

// ---- BR1 interface average / boundary flux at the face order: fFlux[d*5 + q] = u* n_d J_f --------------------------------
struct MxGradFace {
    MixedDev m; Phys ph;
    __device__ void operator()(long long g) const {
        const int f = m.faceNodeFace[g];
        double QL[5], nh[3], UL[5], UR[5];
        for (int q = 0; q < 5; ++q) QL[q] = m.fQ[(long long)q * m.nFaceNodes + g];
        for (int d = 0; d < 3; ++d) nh[d] = m.fN[(long long)d * m.nFaceNodes + g];
        const double Jf = m.fJ[g];
        get_gradients(ph, QL, UL);
        const bool ip = ph.viscous == H3D_VISCOUS_IP;  // IP_GradientInterfaceSolution[Boundary] (EllipticIP.f90:410-585): Uhat = 1/2 (UL - UR) J_f
        if (m.faceType[f] != H3D_FACE_BOUNDARY) {      // interior and MPI faces (BR1_ComputeMPIFaceAverage, EllipticBR1.f90:629-684)
            double QR[5];
            for (int q = 0; q < 5; ++q) QR[q] = m.fQ[(long long)(5 + q) * m.nFaceNodes + g];
            get_gradients(ph, QR, UR);
            for (int q = 0; q < 5; ++q) {
                const double uStar = ip ? 0.5 * (UL[q] - UR[q]) * Jf : 0.5 * (UR[q] - UL[q]) * Jf;
                for (int d = 0; d < 3; ++d) m.fFlux[(long long)(d * 5 + q) * m.nFaceNodes + g] = uStar * nh[d];
            }
        } else if (ip) {                                // the boundary state comes from StateForEqn (FlowState)
            const int zone = m.faceZone[f];
            double Qe[5];
            for (int q = 0; q < 5; ++q) Qe[q] = QL[q];
            bc_flow_state(ph, m.bcType[zone], m.bcParams + 16 * zone, nh, Qe);
            get_gradients(ph, Qe, UR);
            for (int q = 0; q < 5; ++q) {
                const double Uhat = 0.5 * (UL[q] - UR[q]) * Jf;
                for (int d = 0; d < 3; ++d) m.fFlux[(long long)(d * 5 + q) * m.nFaceNodes + g] = Uhat * nh[d];
            }
        } else {
            const int zone = m.faceZone[f];
            for (int q = 0; q < 5; ++q) UR[q] = UL[q];
            bc_grad_vars<true>(ph, m.bcType[zone], m.bcParams + 16 * zone, nh, QL, UR);
            for (int q = 0; q < 5; ++q) for (int d = 0; d < 3; ++d) m.fFlux[(long long)(d * 5 + q) * m.nFaceNodes + g] = (UR[q] - UL[q]) * nh[d] * Jf;
        }
    }
};

// ---- Face_ProjectFluxToElements / ...GradientFluxToElements: nv face fields -> the element side's storage at ITS order -----
struct MxProject {
    MixedDev m; int nv; const double* src; double* dst; double factor;   // src [nv][nFaceNodes], dst [nv][nTrace]; right side *= factor
    __device__ void operator()(long long t) const {
        const int owner = m.traceOwner[t], e = owner / 6, lf = owner % 6;
        const int f = m.elemFace[6 * e + lf], side = m.elemFaceSide[6 * e + lf];
        const int* fo = m.fo + 6 * f;
        const int Nf1 = fo[0], Nf2 = fo[1], Ns1 = fo[2 + 2 * side], Ns2 = fo[3 + 2 * side];
        const int local = (int)(t - m.tOff[owner]);
        int i, j;
        if (side == 0) { i = local % (Ns1 + 1); j = local / (Ns1 + 1); }
        else {
            const int rot = m.faceRot[f];
            const bool swap = rot == 1 || rot == 3 || rot == 4 || rot == 6;
            const int ne1 = (swap ? Ns2 : Ns1) + 1;
            mxRight2Left(local % ne1, local / ne1, Ns1, Ns2, rot, i, j);
        }
        const int pt = m.proj[2 * f + side];
        const long long fb = m.fOff[f];
        const double* T1 = (pt & 1) ? mxT(m, Nf1, Ns1) + i * (Nf1 + 1) : nullptr;
        const double* T2 = (pt & 2) ? mxT(m, Nf2, Ns2) + j * (Nf2 + 1) : nullptr;
        for (int c = 0; c < nv; ++c) {
            const double* F = src + (long long)c * m.nFaceNodes + fb;
            double acc = 0.0;
            if (pt == 0) acc = F[j * (Nf1 + 1) + i];
            else if (pt == 1) { for (int l = 0; l <= Nf1; ++l) acc = acc + T1[l] * F[j * (Nf1 + 1) + l]; }
            else if (pt == 2) { for (int l = 0; l <= Nf2; ++l) acc = acc + T2[l] * F[l * (Nf1 + 1) + i]; }
            else { for (int l = 0; l <= Nf2; ++l) for (int mm = 0; mm <= Nf1; ++mm) acc = acc + T1[mm] * T2[l] * F[l * (Nf1 + 1) + mm]; }
            dst[(long long)c * m.nTrace + t] = side ? factor * acc : acc;
        }
    }
};

// the six element sides of a node in the order of the reference's face integrals: LEFT, RIGHT, FRONT, BACK, BOTTOM, TOP
struct MxSides { long long s[6]; double b[6]; };
__device__ __forceinline__ MxSides mxSides(const MixedDev& m, const MxNode& t) {
    MxSides r;
    const long long* to = m.tOff + 6 * (long long)t.e;
    const double* bx = mxB(m, t.nx - 1); const double* by = mxB(m, t.ny - 1); const double* bz = mxB(m, t.nz - 1);
    r.s[0] = to[5] + t.k * t.ny + t.j; r.b[0] = bx[t.i];            // ELEFT
    r.s[1] = to[3] + t.k * t.ny + t.j; r.b[1] = bx[t.nx + t.i];     // ERIGHT
    r.s[2] = to[0] + t.k * t.nx + t.i; r.b[2] = by[t.j];            // EFRONT
    r.s[3] = to[1] + t.k * t.nx + t.i; r.b[3] = by[t.ny + t.j];     // EBACK
    r.s[4] = to[2] + t.j * t.nx + t.i; r.b[4] = bz[t.k];            // EBOTTOM
    r.s[5] = to[4] + t.j * t.nx + t.i; r.b[5] = bz[t.nz + t.k];     // ETOP
    return r;
}

// ---- BR1_GradientFaceLoop: U_d += (sum over the six sides of unStar_d b) / J ----------------------------------------------
struct MxLift {
    MixedDev m; int ipVariant;   // 0: BR1 (U += faceInt / J); otherwise the interior-penalty variant: U += faceInt * (IPmethod / J) (EllipticIP.f90:366-406)
    int ip;
    __device__ void operator()(long long g) const {
        const MxNode t = mxNode(m, g);
        const MxSides sd = mxSides(m, t);
        const double iJ = ip ? ipVariant * m.invJ[g] : m.invJ[g];
        for (int d = 0; d < 3; ++d) {
            double* Ud = d == 0 ? m.Ux : (d == 1 ? m.Uy : m.Uz);
            for (int q = 0; q < 5; ++q) {
                const double* H = m.unStarE + (long long)(d * 5 + q) * m.nTrace;
                double fi = H[sd.s[0]] * sd.b[0];
                for (int s = 1; s < 6; ++s) fi = fi + H[sd.s[s]] * sd.b[s];
                Ud[(long long)q * m.nNodes + g] = Ud[(long long)q * m.nNodes + g] + fi * iJ;
            }
        }
    }
};

// ---- contravariant fluxes at the nodes: Fc[d*5 + q] = inviscid - viscous ---------------------------------------------------
struct MxFlux {
    MixedDev m; Phys ph; int split;   // split form: Fc = the viscous contravariant flux alone (the inviscid part is formed pairwise in MxVolumeSplit)
    __device__ void operator()(long long g) const {
        double Q[5], ja[9], F[5][3], Fi[15], Fv[15];
        for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * m.nNodes + g];
        for (int q = 0; q < 9; ++q) ja[q] = m.Ja[(long long)q * m.nNodes + g];
        euler_flux(ph, Q, F);
        for (int d = 0; d < 3; ++d) for (int q = 0; q < 5; ++q) Fi[d * 5 + q] = F[q][0] * ja[3 * d] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
        for (int q = 0; q < 15; ++q) Fv[q] = 0.0;
        if (ph.ns) {
            double gx[5], gy[5], gz[5], mu, kappa;
            for (int q = 0; q < 5; ++q) { gx[q] = m.Ux[(long long)q * m.nNodes + g]; gy[q] = m.Uy[(long long)q * m.nNodes + g]; gz[q] = m.Uz[(long long)q * m.nNodes + g]; }
            laminar_mu_kappa(ph, Q, mu, kappa);
            if (ph.les != H3D_LES_NONE) {
                const double mut = smagorinsky<true>(ph, m.lesDelta[m.nodeElem[g]], ph.wallModel ? m.dWall[g] : 0.0, Q, gx, gy, gz);
                mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa;
            }
            viscous_flux<true>(ph, Q, gx, gy, gz, mu, 0.0, kappa, F);
            for (int d = 0; d < 3; ++d) for (int q = 0; q < 5; ++q) Fv[d * 5 + q] = F[q][0] * ja[3 * d] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
        }
        for (int q = 0; q < 15; ++q) m.Fc[(long long)q * m.nNodes + g] = split ? Fv[q] : Fi[q] - Fv[q];
    }
};

// ---- interface and boundary fluxes at the face order: fFlux[q] = (F*_inv - F*_visc) J_f ------------------------------------
struct MxRiemann {
    MixedDev m; Phys ph;
    __device__ void operator()(long long g) const {
        const int f = m.faceNodeFace[g];
        const long long fs = m.nFaceNodes;
        double nh[3], t1[3], t2[3], QL[5], QR[5], visc[5] = {0, 0, 0, 0, 0}, inv[5];
        for (int d = 0; d < 3; ++d) { nh[d] = m.fN[d * fs + g]; t1[d] = m.fT1[d * fs + g]; t2[d] = m.fT2[d * fs + g]; }
        const double Jf = m.fJ[g];
        for (int q = 0; q < 5; ++q) QL[q] = m.fQ[q * fs + g];
        double gx[5], gy[5], gz[5], mu, kappa;
        const bool les = ph.les != H3D_LES_NONE;
        const double dW = (les && ph.wallModel) ? m.fDWall[g] : 0.0;
        if (m.faceType[f] != H3D_FACE_BOUNDARY) {      // interior and MPI faces (computeMPIFaceFlux, SpatialDiscretization.f90:1801-1894)
            for (int q = 0; q < 5; ++q) QR[q] = m.fQ[(5 + q) * fs + g];
            if (ph.ns) {   // BR1_RiemannSolver (EllipticBR1.f90:816-868)
                double FL[5][3], FR[5][3];
                for (int q = 0; q < 5; ++q) { gx[q] = m.fU[(0 * 10 + q) * fs + g]; gy[q] = m.fU[(1 * 10 + q) * fs + g]; gz[q] = m.fU[(2 * 10 + q) * fs + g]; }
                laminar_mu_kappa(ph, QL, mu, kappa);
                if (les) { const double mut = smagorinsky<true>(ph, m.fDelta[f], dW, QL, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                viscous_flux<true>(ph, QL, gx, gy, gz, mu, 0.0, kappa, FL);
                for (int q = 0; q < 5; ++q) { gx[q] = m.fU[(0 * 10 + 5 + q) * fs + g]; gy[q] = m.fU[(1 * 10 + 5 + q) * fs + g]; gz[q] = m.fU[(2 * 10 + 5 + q) * fs + g]; }
                laminar_mu_kappa(ph, QR, mu, kappa);
                if (les) { const double mut = smagorinsky<true>(ph, m.fDelta[f], dW, QR, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                viscous_flux<true>(ph, QR, gx, gy, gz, mu, 0.0, kappa, FR);
                for (int q = 0; q < 5; ++q) {
                    const double fx = 0.5 * (FL[q][0] + FR[q][0]), fy = 0.5 * (FL[q][1] + FR[q][1]), fz = 0.5 * (FL[q][2] + FR[q][2]);
                    visc[q] = fx * nh[0] + fy * nh[1] + fz * nh[2];
                }
                if (ph.viscous == H3D_VISCOUS_IP) {   // IP_RiemannSolver (EllipticIP.f90:704-761)
                    const double penalty = m.fPenalty[f];
                    for (int q = 0; q < 5; ++q) visc[q] = visc[q] - penalty * ph.mu * (QL[q] - QR[q]);
                }
            }
        } else {
            const int zone = m.faceZone[f]; const int btype = m.bcType[zone]; const double* P = m.bcParams + 16 * zone;
            for (int q = 0; q < 5; ++q) QR[q] = QL[q];
            bc_flow_state(ph, btype, P, nh, QR);
            if (ph.ns) {
                double F[5][3];
                for (int q = 0; q < 5; ++q) { gx[q] = m.fU[(0 * 10 + q) * fs + g]; gy[q] = m.fU[(1 * 10 + q) * fs + g]; gz[q] = m.fU[(2 * 10 + q) * fs + g]; }
                laminar_mu_kappa(ph, QL, mu, kappa);
                if (les) { const double mut = smagorinsky<true>(ph, m.fDelta[f], dW, QL, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                viscous_flux<true>(ph, QL, gx, gy, gz, mu, 0.0, kappa, F);
                for (int q = 0; q < 5; ++q) visc[q] = F[q][0] * nh[0] + F[q][1] * nh[1] + F[q][2] * nh[2];
                bc_neumann(btype, P, QL, visc);
            }
        }
        riemann_solver<true>(ph, QL, QR, nh, t1, t2, inv);
        for (int q = 0; q < 5; ++q) m.fFlux[q * fs + g] = (inv[q] - visc[q]) * Jf;
    }
};

// ---- MPI faces: HexMesh_UpdateMPIFacesSolution / ...Gradients (HexMesh.f90:1199-1391): the local side's traces at the face
//      order go to the neighbour, which stores them as ITS remote side (both ranks hold the face frame of the global face).
//      nSets field sets of 5 (Q: 1; gradients: 3, set stride 10 fields).  Message of neighbour b: [set][c][node of b].
struct MxHaloPack {
    MixedDev m; int nSets; const double* src; double* buf; int remote;   // remote = 0: pack the local side; 1: unpack into the remote side
    __device__ void operator()(long long hn) const {
        const int k = m.hNodeFace[hn], f = m.haloFace[k], side = remote ? 1 - m.haloSide[k] : m.haloSide[k], b = m.haloNbr[k];
        const long long g = m.fOff[f] + (hn - m.hOff[k]);
        const long long nb = m.nbrNodeOff[b + 1] - m.nbrNodeOff[b], base = (long long)nSets * 5 * m.nbrNodeOff[b] + (hn - m.nbrNodeOff[b]);
        for (int st = 0; st < nSets; ++st) for (int c = 0; c < 5; ++c) {
            double* fld = const_cast<double*>(src) + (long long)(st * 10 + side * 5 + c) * m.nFaceNodes + g;
            double* msg = buf + base + (long long)(st * 5 + c) * nb;
            if (remote) *fld = *msg; else *msg = *fld;
        }
    }
};

// ---- weak volume integral + surface integral, /J, +S, Runge-Kutta update ---------------------------------------------------
// surface integral of the six sides, /J, +S, and the update of one Runge-Kutta stage for equation q of node g
__device__ __forceinline__ void mxSurfaceAndUpdate(const MixedDev& m, const MxRk& rk, const MxSides& sd, long long g, int q, double vol, double Jn) {
    const double* Fs = m.fStarE + (long long)q * m.nTrace;
    double fi = Fs[sd.s[0]] * sd.b[0];
    for (int s = 1; s < 6; ++s) fi = fi + Fs[sd.s[s]] * sd.b[s];
    double res = vol - fi;
    res = res / Jn;
    if (m.S) res = res + m.S[(long long)q * m.nNodes + g];
    const long long o = (long long)q * m.nNodes + g;
    m.QDot[o] = res;
    if (rk.mode == 1) {
        const double gg = rk.a * m.G[o] + res;
        m.G[o] = gg;
        m.Q[o] = m.Q[o] + rk.cdt * gg;
    } else if (rk.mode == 2) {   // TakeSSPRK33Step / TakeSSPRK43Step
        const double Qk = m.Q[o];
        const double g0 = rk.copyG ? Qk : m.G[o];
        if (rk.copyG) m.G[o] = g0;
        m.Q[o] = rk.a * g0 + rk.b * Qk + rk.cdt * res;
    }
}

struct MxVolume {
    MixedDev m; MxRk rk;
    __device__ void operator()(long long g) const {
        const MxNode t = mxNode(m, g);
        const MxSides sd = mxSides(m, t);
        const double* hx = mxHatD(m, t.nx - 1) + t.i * t.nx; const double* hy = mxHatD(m, t.ny - 1) + t.j * t.ny; const double* hz = mxHatD(m, t.nz - 1) + t.k * t.nz;
        const double Jn = m.J[g];
        for (int q = 0; q < 5; ++q) {
            const double* Fx = m.Fc + (long long)(0 * 5 + q) * m.nNodes + t.base; const double* Fy = m.Fc + (long long)(1 * 5 + q) * m.nNodes + t.base;
            const double* Fz = m.Fc + (long long)(2 * 5 + q) * m.nNodes + t.base;
            double vol = 0.0;
            for (int l = 0; l < t.nx; ++l) vol = vol + hx[l] * Fx[((long long)t.k * t.ny + t.j) * t.nx + l];
            for (int l = 0; l < t.ny; ++l) vol = vol + hy[l] * Fy[((long long)t.k * t.ny + l) * t.nx + t.i];
            for (int l = 0; l < t.nz; ++l) vol = vol + hz[l] * Fz[((long long)l * t.ny + t.j) * t.nx + t.i];
            mxSurfaceAndUpdate(m, rk, sd, g, q, vol, Jn);
        }
    }
};

// ---- split form: SplitDG_ComputeSplitFormFluxes (HyperbolicSplitForm.f90:64-116) + ScalarWeakIntegrals_SplitVolumeDivergence
//      (DGIntegrals.f90:92-129), QDot = -volInt (SpatialDiscretization.f90:1678).  A pair (a < b) of nodes of a line has ONE two-point
//      flux, F#(Q_a, Q_b): both nodes evaluate it with the lower node first, as the reference computes it once and mirrors it.
//      The pair fluxes read the state of OTHER nodes, so the Runge-Kutta update cannot be fused here (a thread would overwrite Q
//      while another still reads it): this functor stores QDot and MxUpdate applies the stage update afterwards.
struct MxUpdate {
    MixedDev m; MxRk rk;
    __device__ void operator()(long long t) const {      // t over 5 nNodes: one value of the state
        const double res = m.QDot[t];
        if (rk.mode == 1) {
            const double gg = rk.a * m.G[t] + res;
            m.G[t] = gg;
            m.Q[t] = m.Q[t] + rk.cdt * gg;
        } else if (rk.mode == 2) {
            const double Qk = m.Q[t];
            const double g0 = rk.copyG ? Qk : m.G[t];
            if (rk.copyG) m.G[t] = g0;
            m.Q[t] = rk.a * g0 + rk.b * Qk + rk.cdt * res;
        }
    }
};
struct MxVolumeSplit {
    MixedDev m; Phys ph;
    __device__ void operator()(long long g) const {
        const MxNode t = mxNode(m, g);
        const MxSides sd = mxSides(m, t);
        const int nn[3] = {t.nx, t.ny, t.nz}, ijk[3] = {t.i, t.j, t.k};
        const long long stride[3] = {1, t.nx, (long long)t.nx * t.ny};
        double Q[5], vol[5] = {0, 0, 0, 0, 0};
        for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * m.nNodes + g];
        for (int d = 0; d < 3; ++d) {
            const int n = nn[d], me = ijk[d];
            const double* sharp = mxSharpD(m, n - 1) + me * n; const double* hat = mxHatD(m, n - 1) + me * n;
            double ja[3];
            for (int c = 0; c < 3; ++c) ja[c] = m.Ja[(long long)(3 * d + c) * m.nNodes + g];
            for (int l = 0; l < n; ++l) {
                const long long g2 = g + (long long)(l - me) * stride[d];
                double fs[5];
                if (l == me) {   // the consistent flux on the diagonal: the contravariant Euler flux of the node
                    double F[5][3];
                    euler_flux(ph, Q, F);
                    for (int q = 0; q < 5; ++q) fs[q] = F[q][0] * ja[0] + F[q][1] * ja[1] + F[q][2] * ja[2];
                } else {
                    double Q2[5], ja2[3];
                    for (int q = 0; q < 5; ++q) Q2[q] = m.Q[(long long)q * m.nNodes + g2];
                    for (int c = 0; c < 3; ++c) ja2[c] = m.Ja[(long long)(3 * d + c) * m.nNodes + g2];
                    if (l > me) two_point_flux<true>(ph, Q, Q2, ja, ja2, fs); else two_point_flux<true>(ph, Q2, Q, ja2, ja, fs);
                }
                for (int q = 0; q < 5; ++q) vol[q] = vol[q] + sharp[l] * fs[q] + hat[l] * m.Fc[(long long)(d * 5 + q) * m.nNodes + g2];
            }
        }
        const double Jn = m.J[g];
        const MxRk none{0, 0.0, 0.0, 0.0, 0};
        for (int q = 0; q < 5; ++q) mxSurfaceAndUpdate(m, none, sd, g, q, -vol[q], Jn);
    }
};

// ---- global2LocalQ / local2GlobalQ (StorageClass.f90:390-440, 549-579): the reference's packed [node][C] <-> the device's [c][node]
struct MxAosToSoa {
    MixedDev m; int C; const double* src; double* dst;
    __device__ void operator()(long long t) const { for (int c = 0; c < C; ++c) dst[(long long)c * m.nNodes + t] = src[t * C + c]; }
};
struct MxSoaToAos {
    MixedDev m; int C; const double* src; double* dst;
    __device__ void operator()(long long t) const { for (int c = 0; c < C; ++c) dst[t * C + c] = src[(long long)c * m.nNodes + t]; }
};

// ---- stage_limiter (ExplicitMethods.f90:1755-1847): one thread per element; density, then pressure, scaled towards the element
//      average so that they stay above min(minimum, average)
struct MxLimiter {
    MixedDev m; Phys ph; double limiterMin;
    __device__ void operator()(long long e) const {
        const int nx = m.eN[3 * e], ny = m.eN[3 * e + 1], nz = m.eN[3 * e + 2];
        const double* wx = mxW(m, nx - 1); const double* wy = mxW(m, ny - 1); const double* wz = mxW(m, nz - 1);
        const long long g0 = m.eOff[e], g1 = m.eOff[e + 1];
        double Qavg[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            const long long g = g0 + ((long long)k * ny + j) * nx + i;
            for (int q = 0; q < 5; ++q) Qavg[q] = Qavg[q] + m.Q[(long long)q * m.nNodes + g] * wx[i] * wy[j] * wz[k] * m.J[g];
        }
        for (int q = 0; q < 5; ++q) Qavg[q] = Qavg[q] / m.volume[e];
        double minrho = 1.7976931348623157e308;
        for (long long g = g0; g < g1; ++g) { const double rho = m.Q[g]; if (rho < minrho) minrho = rho; }
        if (Qavg[0] != minrho) {
            const double mm = fmin(limiterMin, Qavg[0]);
            const double theta = fabs((Qavg[0] - mm) / (Qavg[0] - minrho));
            if (theta <= 1.0) for (long long g = g0; g < g1; ++g) m.Q[g] = theta * (m.Q[g] - Qavg[0]) + Qavg[0];
        }
        double minp = 1.7976931348623157e308, pavg = 0.0;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            const long long g = g0 + ((long long)k * ny + j) * nx + i;
            double Q[5];
            for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * m.nNodes + g];
            const double p = ph.gm1 * (Q[4] - 0.5 * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]) / Q[0]);
            pavg = pavg + p * wx[i] * wy[j] * wz[k] * m.J[g];
            if (p < minp) minp = p;
        }
        pavg = pavg / m.volume[e];
        if (pavg != minp) {
            const double mm = fmin(limiterMin, pavg);
            const double theta = fabs((pavg - mm) / (pavg - minp));
            if (theta <= 1.0) for (long long g = g0; g < g1; ++g) for (int q = 0; q < 5; ++q) {
                double& v = m.Q[(long long)q * m.nNodes + g];
                v = theta * (v - Qavg[q]) + Qavg[q];
            }
        }
    }
};

// ---- StatisticsMonitor_UpdateValues (StatisticsMonitor.f90:279-540): running averages, data [var][nNodes] -----------------------
struct MxStatistics {
    MixedDev m; int nv; double ratio, inv; double* data;
    __device__ void operator()(long long t) const {
        const long long nn = m.nNodes;
        double Q[5];
        for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * nn + t];
        const double r1 = inv / Q[0], r2 = inv / pow2(Q[0]);
        double* d = data + t;
        d[0 * nn] = d[0 * nn] * ratio + Q[1] * r1;
        d[1 * nn] = d[1 * nn] * ratio + Q[2] * r1;
        d[2 * nn] = d[2 * nn] * ratio + Q[3] * r1;
        d[3 * nn] = d[3 * nn] * ratio + pow2(Q[1]) * r2;
        d[4 * nn] = d[4 * nn] * ratio + pow2(Q[2]) * r2;
        d[5 * nn] = d[5 * nn] * ratio + pow2(Q[3]) * r2;
        d[6 * nn] = d[6 * nn] * ratio + Q[1] * Q[2] * r2;
        d[7 * nn] = d[7 * nn] * ratio + Q[1] * Q[3] * r2;
        d[8 * nn] = d[8 * nn] * ratio + Q[2] * Q[3] * r2;
        for (int q = 0; q < 5; ++q) d[(long long)(9 + q) * nn] = d[(long long)(9 + q) * nn] * ratio + Q[q] * inv;
        if (nv == 29) for (int q = 0; q < 5; ++q) {
            d[(long long)(14 + q) * nn] = d[(long long)(14 + q) * nn] * ratio + m.Ux[(long long)q * nn + t] * inv;
            d[(long long)(19 + q) * nn] = d[(long long)(19 + q) * nn] * ratio + m.Uy[(long long)q * nn + t] * inv;
            d[(long long)(24 + q) * nn] = d[(long long)(24 + q) * nn] * ratio + m.Uz[(long long)q * nn + t] * inv;
        }
    }
};

// ---- reductions: one thread per element (face), nodes in the reference's order; partial[q][e] ---------------------------------
struct MxRedResidual {   // ComputeMaxResiduals (DGSEMClass.f90:770-856) + checkForNan on Q
    MixedDev m;
    __device__ void operator()(long long e) const {
        double v[6] = {0, 0, 0, 0, 0, 0};
        for (long long g = m.eOff[e]; g < m.eOff[e + 1]; ++g) for (int q = 0; q < 5; ++q) {
            v[q] = fmax(v[q], fabs(m.QDot[(long long)q * m.nNodes + g]));
            if (isnan(m.Q[(long long)q * m.nNodes + g])) v[5] = 1.0;
        }
        for (int q = 0; q < 6; ++q) m.partial[q * m.partialStride + e] = v[q];
    }
};
struct MxRedTimestep {   // MaxTimeStep (DGSEMClass.f90:870-1034): the spacings of the element's own nodal storages
    MixedDev m; Phys ph; double cfl, dcfl;
    __device__ void operator()(long long e) const {
        double dxi[3];
        for (int d = 0; d < 3; ++d) { const int N = m.eN[3 * e + d] - 1; const double* x = mxX(m, N); dxi[d] = N != 0 ? 1.0 / fabs(x[1] - x[0]) : 0.0; }
        double vc = 1.7976931348623157e308, vv = 1.7976931348623157e308;
        for (long long g = m.eOff[e]; g < m.eOff[e + 1]; ++g) {
            double Q[5], ja[9];
            for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * m.nNodes + g];
            for (int c = 0; c < 9; ++c) ja[c] = m.Ja[(long long)c * m.nNodes + g];
            const double u = fabs(Q[1] / Q[0]), v = fabs(Q[2] / Q[0]), w = fabs(Q[3] / Q[0]);
            const double p = pressure(ph, Q);
            const double a = sqrt(ph.gamma * p / Q[0]);
            const double e0 = u + a, e1 = v + a, e2 = w + a;
            const double jac = m.J[g];
            const double l1 = fabs(ja[0] * e0 + ja[1] * e1 + ja[2] * e2) * dxi[0];
            const double l2 = fabs(ja[3] * e0 + ja[4] * e1 + ja[5] * e2) * dxi[1];
            const double l3 = fabs(ja[6] * e0 + ja[7] * e1 + ja[8] * e2) * dxi[2];
            vc = fmin(vc, cfl * fabs(jac) / (l1 + l2 + l3));
            if (ph.ns) {
                const double T = ph.gammaM2 * p / Q[0];
                const double mu = sutherland(ph, T);
                const double v1 = mu * (dxi[0] * dxi[0]) * fabs(ja[0] + ja[1] + ja[2]);
                const double v2 = mu * (dxi[1] * dxi[1]) * fabs(ja[3] + ja[4] + ja[5]);
                const double v3 = mu * (dxi[2] * dxi[2]) * fabs(ja[6] + ja[7] + ja[8]);
                vv = fmin(vv, dcfl * fabs(jac) / (v1 + v2 + v3));
            }
        }
        m.partial[e] = vc; m.partial[m.partialStride + e] = vv;
    }
};
struct MxRedIntegral {   // ScalarVolumeIntegral (VolumeIntegrals.f90:76-120, 167-286), the kinds that need Q, QDot and gradients only
    MixedDev m; Phys ph; int kind;
    __device__ void operator()(long long e) const {
        const int nx = m.eN[3 * e], ny = m.eN[3 * e + 1], nz = m.eN[3 * e + 2];
        const double* wx = mxW(m, nx - 1); const double* wy = mxW(m, ny - 1); const double* wz = mxW(m, nz - 1);
        double loc = 0.0;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            const long long g = m.eOff[e] + ((long long)k * ny + j) * nx + i;
            double Q[5], QD[5];
            for (int q = 0; q < 5; ++q) { Q[q] = m.Q[(long long)q * m.nNodes + g]; QD[q] = m.QDot[(long long)q * m.nNodes + g]; }
            const double wJ = wx[i] * wy[j] * wz[k] * m.J[g];
            switch (kind) {
                case H3D_INT_VOLUME: loc = loc + wJ; break;
                case H3D_INT_KINETIC_ENERGY: {
                    double KinEn = pow2(Q[1]); KinEn = KinEn + pow2(Q[2]); KinEn = KinEn + pow2(Q[3]);
                    KinEn = 0.5 * KinEn / Q[0];
                    loc = loc + wJ * KinEn;
                } break;
                case H3D_INT_KINETIC_ENERGY_RATE: {
                    double uvw = Q[1] / Q[0];
                    double KinEn = uvw * QD[1] - 0.5 * pow2(uvw) * QD[0];
                    uvw = Q[2] / Q[0]; KinEn = KinEn + uvw * QD[2] - 0.5 * pow2(uvw) * QD[0];
                    uvw = Q[3] / Q[0]; KinEn = KinEn + uvw * QD[3] - 0.5 * pow2(uvw) * QD[0];
                    loc = loc + wJ * KinEn;
                } break;
                case H3D_INT_ENSTROPHY: {
                    double gx[5], gy[5], gz[5], U_x[3], U_y[3], U_z[3];
                    for (int q = 0; q < 5; ++q) { gx[q] = m.Ux[(long long)q * m.nNodes + g]; gy[q] = m.Uy[(long long)q * m.nNodes + g]; gz[q] = m.Uz[(long long)q * m.nNodes + g]; }
                    velocity_gradients_gv<true>(ph, Q, gx, gy, gz, U_x, U_y, U_z);
                    const double KinEn = pow2(U_y[2] - U_z[1]) + pow2(U_z[0] - U_x[2]) + pow2(U_x[1] - U_y[0]);
                    loc = loc + wJ * KinEn;
                } break;
                case H3D_INT_VELOCITY: loc = loc + wx[i] * wy[j] * wz[k] * sqrt(pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) / Q[0] * m.J[g]; break;
                case H3D_INT_INTERNAL_ENERGY: loc = loc + wJ * Q[4]; break;
                case H3D_INT_ENTROPY: { const double pr = pressure(ph, Q); const double sp = log(pr) - ph.gamma * log(Q[0]); loc = loc + wJ * sp; } break;
                case H3D_INT_MATH_ENTROPY: {
                    const double pr = pressure(ph, Q); const double sp = log(pr) - ph.gamma * log(Q[0]);
                    const double ms = -Q[0] * sp / ph.gm1;
                    loc = loc + wJ * ms;
                } break;
                case H3D_INT_ENTROPY_RATE: case H3D_INT_ENTROPY_BALANCE: {      // NSGradientVariables_ENTROPY whatever the gradient variables of the run
                    Phys pe = ph; pe.gradVars = H3D_GRADVARS_ENTROPY;
                    double EV[5];
                    get_gradients(pe, Q, EV);
                    double dot = 0.0;
                    for (int q = 0; q < 5; ++q) dot = dot + QD[q] * EV[q];
                    if (kind == H3D_INT_ENTROPY_BALANCE) {      // + the viscous work with the run's gradient variables (:335-370)
                        double gx[5], gy[5], gz[5], F[5][3], mu, kappa, work = 0.0;
                        for (int q = 0; q < 5; ++q) { gx[q] = m.Ux[(long long)q * m.nNodes + g]; gy[q] = m.Uy[(long long)q * m.nNodes + g]; gz[q] = m.Uz[(long long)q * m.nNodes + g]; }
                        laminar_mu_kappa(ph, Q, mu, kappa);
                        if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<true>(ph, m.lesDelta[e], ph.wallModel ? m.dWall[g] : 0.0, Q, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                        viscous_flux<true>(ph, Q, gx, gy, gz, mu, 0.0, kappa, F);
                        for (int q = 0; q < 5; ++q) work = work + (F[q][0] * gx[q] + F[q][1] * gy[q] + F[q][2] * gz[q]);
                        dot = dot + work;
                    }
                    loc = loc + wJ * dot;
                } break;
                case H3D_INT_KINETIC_ENERGY_BALANCE: {      // kinetic energy rate + viscous work - pressure work + de-aliasing correction (:220-265, 724-764)
                    const int nn3[3] = {nx, ny, nz}, ijk[3] = {i, j, k};
                    const long long stride[3] = {1, nx, (long long)nx * ny};
                    double gMp[3] = {0, 0, 0}, Mgp[3] = {0, 0, 0};
                    for (int ax = 0; ax < 3; ++ax) {
                        const double* Dr = mxD(m, nn3[ax] - 1) + ijk[ax] * nn3[ax];
                        for (int l = 0; l < nn3[ax]; ++l) {
                            const long long gl = g + (long long)(l - ijk[ax]) * stride[ax];
                            double Ql[5];
                            for (int q = 0; q < 5; ++q) Ql[q] = m.Q[(long long)q * m.nNodes + gl];
                            const double pl = pressure(ph, Ql);
                            for (int c = 0; c < 3; ++c) {
                                gMp[c] = gMp[c] + pl * m.Ja[(long long)(3 * ax + c) * m.nNodes + gl] * Dr[l];
                                Mgp[c] = Mgp[c] + pl * m.Ja[(long long)(3 * ax + c) * m.nNodes + g] * Dr[l];
                            }
                        }
                    }
                    double gx[5], gy[5], gz[5], F[5][3], mu, kappa;
                    for (int q = 0; q < 5; ++q) { gx[q] = m.Ux[(long long)q * m.nNodes + g]; gy[q] = m.Uy[(long long)q * m.nNodes + g]; gz[q] = m.Uz[(long long)q * m.nNodes + g]; }
                    const double inv_rho = 1.0 / Q[0];
                    double uvw = Q[1] * inv_rho;
                    double ke = uvw * QD[1] - 0.5 * pow2(uvw) * QD[0];
                    uvw = Q[2] * inv_rho; ke = ke + uvw * QD[2] - 0.5 * pow2(uvw) * QD[0];
                    uvw = Q[3] * inv_rho; ke = ke + uvw * QD[3] - 0.5 * pow2(uvw) * QD[0];
                    const double p3 = ph.gm1 * (Q[4] - 0.5 * (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) * inv_rho);
                    const double corr = 0.5 * (Q[1] * (Mgp[0] - gMp[0]) + Q[2] * (Mgp[1] - gMp[1]) + Q[3] * (Mgp[2] - gMp[2])) * inv_rho;
                    Phys pE = ph; pE.gradVars = H3D_GRADVARS_ENERGY;
                    laminar_mu_kappa(ph, Q, mu, kappa);
                    if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<true>(ph, m.lesDelta[e], ph.wallModel ? m.dWall[g] : 0.0, Q, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                    viscous_flux<true>(pE, Q, gx, gy, gz, mu, 0.0, kappa, F);
                    double work = 0.0;
                    for (int q = 1; q < 4; ++q) work = work + (F[q][0] * gx[q] + F[q][1] * gy[q] + F[q][2] * gz[q]);
                    loc = loc + wx[i] * wy[j] * wz[k] * (m.J[g] * (ke + work - p3 * (gx[1] + gy[2] + gz[3])) + corr);
                } break;
                default: break;
            }
        }
        m.partial[e] = loc;
    }
};
struct MxRedSurface {   // ScalarSurfaceIntegral / VectorSurfaceIntegral (SurfaceIntegrals.f90:40-445) on the faces of one zone
    MixedDev m; Phys ph; int zone, kind;
    __device__ void operator()(long long f) const {
        double fv[3] = {0, 0, 0};
        if (m.faceType[f] == H3D_FACE_BOUNDARY && m.faceZone[f] == zone) {
            const int n1 = m.fo[6 * f] + 1, n2 = m.fo[6 * f + 1] + 1;
            const double* w1 = mxW(m, n1 - 1); const double* w2 = mxW(m, n2 - 1);
            const long long fs = m.nFaceNodes;
            for (int j = 0; j < n2; ++j) for (int i = 0; i < n1; ++i) {
                const long long g = m.fOff[f] + (long long)j * n1 + i;
                double Q[5], nh[3];
                for (int q = 0; q < 5; ++q) Q[q] = m.fQ[q * fs + g];
                for (int d = 0; d < 3; ++d) nh[d] = m.fN[d * fs + g];
                const double Jf = m.fJ[g];
                switch (kind) {
                    case H3D_SURF_SURFACE: fv[0] = fv[0] + w1[i] * w2[j] * Jf; break;
                    case H3D_SURF_MASS_FLOW: fv[0] = fv[0] + (Q[1] * nh[0] + Q[2] * nh[1] + Q[3] * nh[2]) * w1[i] * w2[j] * Jf; break;
                    case H3D_SURF_FLOW_RATE: fv[0] = fv[0] + (1.0 / Q[0]) * (Q[1] * nh[0] + Q[2] * nh[1] + Q[3] * nh[2]) * w1[i] * w2[j] * Jf; break;
                    case H3D_SURF_PRESSURE: { const double pr = pressure(ph, Q); fv[0] = fv[0] + pr * w1[i] * w2[j] * Jf; } break;
                    case H3D_SURF_VEC_SURFACE: for (int d = 0; d < 3; ++d) fv[d] = fv[d] + w1[i] * w2[j] * Jf * nh[d]; break;
                    case H3D_SURF_PRESSURE_FORCE: { const double pr = pressure(ph, Q); for (int d = 0; d < 3; ++d) fv[d] = fv[d] + (pr * nh[d]) * Jf * w1[i] * w2[j]; } break;
                    default: {   // total / viscous force: getStressTensor (Physics_NS.f90:822-886)
                        double gx[5], gy[5], gz[5], U_x[3], U_y[3], U_z[3], tau[3][3], mu, kappa;
                        for (int q = 0; q < 5; ++q) { gx[q] = m.fU[(0 * 10 + q) * fs + g]; gy[q] = m.fU[(1 * 10 + q) * fs + g]; gz[q] = m.fU[(2 * 10 + q) * fs + g]; }
                        stress_velocity_gradients(ph, Q, gx, gy, gz, U_x, U_y, U_z);
                        laminar_mu_kappa(ph, Q, mu, kappa);
                        const double divV = U_x[0] + U_y[1] + U_z[2];
                        tau[0][0] = mu * (2.0 * U_x[0] - 2.0 / 3.0 * divV);
                        tau[1][0] = mu * (U_x[1] + U_y[0]);
                        tau[2][0] = mu * (U_x[2] + U_z[0]);
                        tau[0][1] = tau[1][0];
                        tau[1][1] = mu * (2.0 * U_y[1] - 2.0 / 3.0 * divV);
                        tau[2][1] = mu * (U_y[2] + U_z[1]);
                        tau[0][2] = tau[2][0];
                        tau[1][2] = tau[2][1];
                        tau[2][2] = mu * (2.0 * U_z[2] - 2.0 / 3.0 * divV);
                        const double pr = pressure(ph, Q);
                        for (int d = 0; d < 3; ++d) {
                            const double tn = tau[d][0] * nh[0] + tau[d][1] * nh[1] + tau[d][2] * nh[2];
                            if (kind == H3D_SURF_TOTAL_FORCE) fv[d] = fv[d] + (pr * nh[d] - tn) * Jf * w1[i] * w2[j];
                            else fv[d] = fv[d] - tn * Jf * w1[i] * w2[j];
                        }
                    }
                }
            }
        }
        for (int d = 0; d < 3; ++d) m.partial[d * m.partialStride + f] = fv[d];
    }
};
struct MxProbe {   // Probe_Update (Probe.f90:330-420); Lagrange vectors padded to ld values per direction
    MixedDev m; Phys ph; const int* elem; const int* variable; const double* L; int nProbes, ld; double* values;
    __device__ void operator()(long long pr) const {
        const int e = elem[pr], nx = m.eN[3 * e], ny = m.eN[3 * e + 1], nz = m.eN[3 * e + 2];
        const double* lx = L + (long long)pr * ld; const double* ly = L + (long long)(nProbes + pr) * ld; const double* lz = L + (long long)(2 * nProbes + pr) * ld;
        double value = 0.0;
        for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            const long long g = m.eOff[e] + ((long long)k * ny + j) * nx + i;
            double Q[5], var;
            for (int q = 0; q < 5; ++q) Q[q] = m.Q[(long long)q * m.nNodes + g];
            switch (variable[pr]) {
                case H3D_PROBE_PRESSURE: var = pressure(ph, Q); break;
                case H3D_PROBE_VELOCITY: var = sqrt(pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) / Q[0]; break;
                case H3D_PROBE_U: var = Q[1] / Q[0]; break;
                case H3D_PROBE_V: var = Q[2] / Q[0]; break;
                case H3D_PROBE_W: var = Q[3] / Q[0]; break;
                case H3D_PROBE_MACH: {
                    var = pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3]) / pow2(Q[0]);
                    var = sqrt(var / (ph.gamma * (ph.gamma - 1.0) * (Q[4] / Q[0] - 0.5 * var)));
                } break;
                default: var = 0.5 * (pow2(Q[1]) + pow2(Q[2]) + pow2(Q[3])) / Q[0]; break;   // H3D_PROBE_K
            }
            value = value + var * lx[i] * ly[j] * lz[k];
        }
        values[pr] = value;
    }
};

// ============================================================================================================================
//  Orchestration, shared by the CUDA backend (libh3dgpu.so) and the host-loop backend of tests/emu.
//  Backend B:  template <class T> T* alloc(size_t);  void zero(T*, size_t);  void upload(T* dst, const T* src, size_t);  void download(T* dst, const T* src, size_t)
//              (both synchronous);  template <class F> void launch(const F&, long long count);  const char* error()  (nullptr = ok)
//              void exchange(const double* send, double* recv, int nNbr, const int* ranks, const long long* offset, const long long* count)
//              (device buffers, doubles; in order with the launches);  void allreduce(double* hostValues, int n, int op)  (0 max, 1 min, 2 sum)
// ============================================================================================================================
struct MxBasis { int N = -1, nodeType = 0; std::vector<double> x, w, D, hatD, sharpD, v, b; };

inline bool mxRkCoefficients(int scheme, int k, MxRk& rk, double dt) {
    static const double A3[3] = {0.0, -5.0 / 9.0, -153.0 / 128.0}, C3[3] = {1.0 / 3.0, 15.0 / 16.0, 8.0 / 15.0};
    static const double A5[5] = {0.0, -0.4178904745, -1.192151694643, -1.697784692471, -1.514183444257};
    static const double C5[5] = {0.1496590219993, 0.3792103129999, 0.8229550293869, 0.6994504559488, 0.1530572479681};
    static const double A14[14] = {0.0000000000000000, -0.7188012108672410, -0.7785331173421570, -0.0053282796654044, -0.8552979934029281, -3.9564138245774565, -1.5780575380587385,
                                   -2.0837094552574054, -0.7483334182761610, -0.7032861106563359, +0.0013917096117681, -0.0932075369637460, -0.9514200470875948, -7.1151571693922548};
    static const double C14[14] = {0.0367762454319673, 0.3136296607553959, 0.1531848691869027, 0.0030097086818182, 0.3326293790646110, 0.2440251405350864, 0.3718879239592277,
                                   0.6204126221582444, 0.1524043173028741, 0.0760894927419266, 0.0077604214040978, 0.0024647284755382, 0.0780348340049386, 5.5059777270269628};
    static const double S33A[3] = {1.0, 3.0 / 4.0, 1.0 / 3.0}, S33B[3] = {0.0, 1.0 / 4.0, 2.0 / 3.0}, S33C[3] = {1.0, 1.0 / 4.0, 2.0 / 3.0};
    static const double S43A[4] = {1.0, 0.0, 2.0 / 3.0, 0.0}, S43B[4] = {0.0, 1.0, 1.0 / 3.0, 1.0}, S43C[4] = {0.5, 0.5, 1.0 / 6.0, 0.5};
    switch (scheme) {
        case H3D_EULER: if (k >= 1) return false; rk = MxRk{1, 0.0, dt, 0.0, 0}; return true;
        case H3D_RK3: if (k >= 3) return false; rk = MxRk{1, A3[k], C3[k] * dt, 0.0, 0}; return true;
        case H3D_RK5: if (k >= 5) return false; rk = MxRk{1, A5[k], C5[k] * dt, 0.0, 0}; return true;
        case H3D_LSERK14_4: if (k >= 14) return false; rk = MxRk{1, A14[k], C14[k] * dt, 0.0, 0}; return true;
        case H3D_SSPRK33: if (k >= 3) return false; rk = MxRk{2, S33A[k], S33C[k] * dt, S33B[k], k == 0 ? 1 : 0}; return true;
        case H3D_SSPRK43: if (k >= 4) return false; rk = MxRk{2, S43A[k], S43C[k] * dt, S43B[k], k == 0 ? 1 : 0}; return true;
        default: return false;
    }
}
inline int mxRkStages(int scheme) {
    switch (scheme) { case H3D_EULER: return 1; case H3D_RK3: return 3; case H3D_RK5: return 5; case H3D_LSERK14_4: return 14; case H3D_SSPRK33: return 3; case H3D_SSPRK43: return 4; default: return 0; }
}

template <class B>
struct MixedSolver {
    B be;
    MixedDev m{};
    Phys ph{};
    std::string err;
    std::map<int, MxBasis> sp;                                   // NodalStorage(N)
    std::map<std::pair<int, int>, std::vector<double>> T;        // Tset(Norigin, Ndest)
    bool haveMesh = false, haveBC = false, haveHalo = false, splitForm = false, lesWallModel = false, limited = false, interiorPenalty = false;
    double limiterMin = 1e-13;                 // LIMITED, LIMITER_MIN (ExplicitMethods.f90:28-29)
    double* dStats = nullptr; int statVars = 0, statSamples = 0;
    int nBoundaryFaces = 0, maxZone = -1, nZones = 0, maxNodes1D = 0, nMpiFaces = 0, nranks = 1;
    std::vector<int> nbrRank; std::vector<long long> nbrNodeOff;   // neighbours and their halo-node ranges (host copies)
    long long launches = 0;
    std::vector<long long> hEOff, hFOff;
    std::vector<int> hFo, hFaceType, hFaceElem;
    std::vector<double> hPartial, hBuf;
    double* dSource = nullptr; double* stage = nullptr;
    int* dProbeI = nullptr; double* dProbeD = nullptr; size_t probeCap = 0;

    explicit MixedSolver(const B& b) : be(b) {}

    int fail(const std::string& s) { err = s; return 1; }
    int check() { const char* e = be.error(); if (e) { err = e; return 2; } return 0; }
    template <class F> void launch(const F& f, long long n) { if (n > 0) { be.launch(f, n); ++launches; } }

    void setBasis(int N, int nodeType, const double* x, const double* w, const double* D, const double* hatD, const double* sharpD, const double* v, const double* b) {
        MxBasis& s = sp[N]; const int n = N + 1;
        s.N = N; s.nodeType = nodeType; s.x.assign(x, x + n); s.w.assign(w, w + n); s.D.assign(D, D + n * n); s.hatD.assign(hatD, hatD + n * n);
        s.sharpD.assign(n * n, 0.0); if (sharpD) s.sharpD.assign(sharpD, sharpD + n * n);
        s.v.assign(v, v + 2 * n); s.b.assign(b, b + 2 * n);
    }
    int setInterpolation(int No, int Nd, const double* Tm) {
        if (No < 0 || Nd < 0 || No >= MX_MAXN || Nd >= MX_MAXN) return fail("h3d_set_interpolation: polynomial order out of range");
        T[{No, Nd}].assign(Tm, Tm + (size_t)(No + 1) * (Nd + 1));
        return 0;
    }
    template <class U> int up(const std::vector<U>& src, const U** dst) {
        U* d = be.template alloc<U>(std::max<size_t>(src.size(), 1));
        if (!d) return check() ? 2 : fail("device allocation failed");
        if (!src.empty()) be.upload(d, src.data(), src.size());
        *dst = d;
        return check();
    }
    int field(double** dst, size_t count) {
        double* d = be.template alloc<double>(std::max<size_t>(count, 1));
        if (!d) return check() ? 2 : fail("device allocation failed");
        if (count) be.zero(d, count);
        *dst = d;
        return check();
    }
    // [node][C] (the reference's packed order) -> [c][node]
    static void toSoA(const double* src, size_t nn, int C, std::vector<double>& out) {
        out.resize(nn * C);
        for (size_t g = 0; g < nn; ++g) for (int c = 0; c < C; ++c) out[(size_t)c * nn + g] = src[g * C + c];
    }

    // faceOrder (may be null on a single rank): f % Nf, NfLeft, NfRight of every face, [nFace][6]; needed for MPI faces, whose remote
    // element is not in this partition (the reference exchanges it, HexMesh_UpdateMPIFacesPolynomial)
    int setMesh(const H3dPhysics& physics, int nElem, int nFace, const int* elemOrder, const int* faceOrder, const int* elemFace, const int* elemFaceSide, const int* faceElem,
                const int* faceElemSide, const int* faceRot, const int* faceType, const int* faceZone, const double* jGradXi, const double* jGradEta,
                const double* jGradZeta, const double* jacobian, const double* volume, const double* faceNormal, const double* faceT1, const double* faceT2,
                const double* faceJacobian, const double* faceSurface) {
        splitForm = physics.inviscid == H3D_SPLIT_DG;
        if (physics.les != H3D_LES_NONE && (!volume || !faceSurface)) return fail("LES needs the element volumes and face surfaces (h3d_set_mesh_p volume / faceSurface): the filter width would be zero");
        if (physics.flowIsNavierStokes && physics.viscous == H3D_VISCOUS_BR2) return fail("p-nonconforming meshes: BR1 and the interior penalty are the viscous discretizations available (not BR2)");
        interiorPenalty = physics.flowIsNavierStokes && physics.viscous == H3D_VISCOUS_IP;
        if (nElem < 1 || nFace < 1) return fail("h3d_set_mesh_p: empty mesh");
        m.nElem = nElem; m.nFace = nFace;
        std::vector<int> eN(3 * (size_t)nElem), nodeElem, traceOwner, faceNodeFace, fo(6 * (size_t)nFace), proj(2 * (size_t)nFace);
        std::vector<long long> eOff(nElem + 1, 0), tOff(6 * (size_t)nElem + 1, 0), fOff(nFace + 1, 0);
        for (int e = 0; e < nElem; ++e) {
            for (int d = 0; d < 3; ++d) {
                const int N = elemOrder[3 * e + d];
                if (N < 1 || N >= MX_MAXN) return fail("h3d_set_mesh_p: polynomial orders must lie in 1..15");
                if (!sp.count(N)) return fail("h3d_set_mesh_p: h3d_set_basis has not been called for every polynomial order of the mesh");
                if (splitForm && sp[N].nodeType != H3D_GAUSSLOBATTO) return fail("split-form discretization needs Gauss-Lobatto nodes");
                eN[3 * e + d] = N + 1; maxNodes1D = std::max(maxNodes1D, N + 1);
            }
            eOff[e + 1] = eOff[e] + (long long)eN[3 * e] * eN[3 * e + 1] * eN[3 * e + 2];
            static const int ax[6][2] = {{0, 2}, {0, 2}, {0, 1}, {1, 2}, {0, 1}, {1, 2}};
            for (int lf = 0; lf < 6; ++lf) tOff[6 * e + lf + 1] = tOff[6 * e + lf] + (long long)eN[3 * e + ax[lf][0]] * eN[3 * e + ax[lf][1]];
        }
        nBoundaryFaces = 0; maxZone = -1; nMpiFaces = 0;
        for (int f = 0; f < nFace; ++f) {   // Face_LinkWithElements (FaceClass.f90:187-282)
            if (faceType[f] != H3D_FACE_INTERIOR && faceType[f] != H3D_FACE_BOUNDARY && faceType[f] != H3D_FACE_MPI) return fail("h3d_set_mesh_p: unknown face type");
            static const int ax[6][2] = {{0, 2}, {0, 2}, {0, 1}, {1, 2}, {0, 1}, {1, 2}};
            const int rot = faceRot[f];
            if (rot < 0 || rot > 7) return fail("h3d_set_mesh_p: face rotation out of range");
            const bool swapR = rot == 1 || rot == 3 || rot == 4 || rot == 6;
            int* o = &fo[6 * (size_t)f];
            // the orders of a side from its element: side 0 as they are, side 1 in the face frame (swapped for the odd rotations)
            auto sideOrders = [&](int side, int out[2]) -> bool {
                const int e = faceElem[2 * f + side], lf = faceElemSide[2 * f + side];
                if (e < 0 || e >= nElem || lf < 0 || lf > 5) return false;
                const int a = elemOrder[3 * e + ax[lf][0]], b = elemOrder[3 * e + ax[lf][1]];
                out[0] = (side && swapR) ? b : a; out[1] = (side && swapR) ? a : b;
                return true;
            };
            int NfL[2], NfR[2];
            if (faceType[f] == H3D_FACE_MPI) {
                if (nranks < 2) return fail("h3d_set_mesh_p: MPI faces on a single rank");
                if (!faceOrder) return fail("h3d_set_mesh_p: a mesh with MPI faces needs the face orders (faceOrder): the remote element is not in this partition");
                ++nMpiFaces;
                const bool hasL = faceElem[2 * f] >= 0, hasR = faceElem[2 * f + 1] >= 0;
                if (hasL == hasR) return fail("h3d_set_mesh_p: an MPI face has exactly one local side");
                for (int q = 0; q < 6; ++q) o[q] = faceOrder[6 * (size_t)f + q];
                int own[2];
                if (!sideOrders(hasL ? 0 : 1, own) || own[0] != o[hasL ? 2 : 4] || own[1] != o[hasL ? 3 : 5]) return fail("h3d_set_mesh_p: faceOrder of an MPI face contradicts the order of its local element");
                for (int q = 0; q < 6; ++q) if (o[q] < 1 || o[q] >= MX_MAXN) return fail("h3d_set_mesh_p: face orders must lie in 1..15");
                if (o[0] != std::max(o[2], o[4]) || o[1] != std::max(o[3], o[5])) return fail("h3d_set_mesh_p: the order of a face is the maximum of its two sides");
            } else {
                if (!sideOrders(0, NfL)) return fail("h3d_set_mesh_p: face without a left element");
                NfR[0] = NfL[0]; NfR[1] = NfL[1];                         // boundary: NelRight = NelLeft (HexMesh.f90:2470-2472)
                if (faceType[f] == H3D_FACE_INTERIOR) { if (!sideOrders(1, NfR)) return fail("h3d_set_mesh_p: interior face without a right element"); }
                else { ++nBoundaryFaces; maxZone = std::max(maxZone, faceZone[f]); if (faceZone[f] < 0) return fail("h3d_set_mesh_p: boundary face without a zone"); }
                o[0] = std::max(NfL[0], NfR[0]); o[1] = std::max(NfL[1], NfR[1]); o[2] = NfL[0]; o[3] = NfL[1]; o[4] = NfR[0]; o[5] = NfR[1];
                if (faceOrder) for (int q = 0; q < 6; ++q) if (faceOrder[6 * (size_t)f + q] != o[q]) return fail("h3d_set_mesh_p: faceOrder contradicts the orders of the elements");
            }
            for (int s = 0; s < 2; ++s) {
                proj[2 * f + s] = (o[2 + 2 * s] != o[0] ? 1 : 0) + (o[3 + 2 * s] != o[1] ? 2 : 0);
                for (int d = 0; d < 2; ++d) if (o[2 + 2 * s + d] != o[d] && (!T.count({o[2 + 2 * s + d], o[d]}) || !T.count({o[d], o[2 + 2 * s + d]})))
                    return fail("h3d_set_mesh_p: h3d_set_interpolation has not been called for every pair of orders that meet at a face");
            }
            if (!sp.count(o[0]) || !sp.count(o[1])) return fail("h3d_set_mesh_p: h3d_set_basis has not been called for every face order");
            fOff[f + 1] = fOff[f] + (long long)(o[0] + 1) * (o[1] + 1);
        }
        for (int e = 0; e < nElem; ++e) for (int lf = 0; lf < 6; ++lf) {   // e % faceIDs / faceSide must mirror f % elementIDs / elementSide
            const int f = elemFace[6 * e + lf], sd = elemFaceSide[6 * e + lf];
            if (f < 0 || f >= nFace || sd < 0 || sd > 1 || faceElem[2 * f + sd] != e || faceElemSide[2 * f + sd] != lf)
                return fail("h3d_set_mesh_p: elemFace / elemFaceSide do not mirror faceElem / faceElemSide");
        }
        m.nNodes = eOff[nElem]; m.nTrace = tOff[6 * (size_t)nElem]; m.nFaceNodes = fOff[nFace];
        nodeElem.resize(m.nNodes); traceOwner.resize(m.nTrace); faceNodeFace.resize(m.nFaceNodes);
        for (int e = 0; e < nElem; ++e) for (long long g = eOff[e]; g < eOff[e + 1]; ++g) nodeElem[g] = e;
        for (int s = 0; s < 6 * nElem; ++s) for (long long g = tOff[s]; g < tOff[s + 1]; ++g) traceOwner[g] = s;
        for (int f = 0; f < nFace; ++f) for (long long g = fOff[f]; g < fOff[f + 1]; ++g) faceNodeFace[g] = f;
        hEOff = eOff; hFOff = fOff; hFo = fo; hFaceType.assign(faceType, faceType + nFace); hFaceElem.assign(faceElem, faceElem + 2 * (size_t)nFace);
        m.nHalo = 0; m.nNbr = 0; m.nHaloNodes = 0;
        // operators and interpolation matrices
        std::vector<double> ops, ts;
        for (int N = 0; N < MX_MAXN; ++N) {
            m.opBase[N] = -1;
            auto it = sp.find(N);
            if (it == sp.end()) continue;
            m.opBase[N] = (int)ops.size();
            const MxBasis& s = it->second;
            ops.insert(ops.end(), s.D.begin(), s.D.end()); ops.insert(ops.end(), s.hatD.begin(), s.hatD.end());
            ops.insert(ops.end(), s.v.begin(), s.v.end()); ops.insert(ops.end(), s.b.begin(), s.b.end());
            ops.insert(ops.end(), s.w.begin(), s.w.end()); ops.insert(ops.end(), s.x.begin(), s.x.end());
            ops.insert(ops.end(), s.sharpD.begin(), s.sharpD.end());
        }
        for (int a = 0; a < MX_MAXN; ++a) for (int b = 0; b < MX_MAXN; ++b) {
            m.tBase[a][b] = -1;
            auto it = T.find({a, b});
            if (it == T.end()) continue;
            m.tBase[a][b] = (int)ts.size();
            ts.insert(ts.end(), it->second.begin(), it->second.end());
        }
        std::vector<int> vElemFace(elemFace, elemFace + 6 * (size_t)nElem), vElemFaceSide(elemFaceSide, elemFaceSide + 6 * (size_t)nElem);
        std::vector<int> vFaceElem(faceElem, faceElem + 2 * (size_t)nFace), vFaceElemSide(faceElemSide, faceElemSide + 2 * (size_t)nFace);
        std::vector<int> vRot(faceRot, faceRot + nFace), vType(faceType, faceType + nFace), vZone(faceZone, faceZone + nFace);
        if (up(eOff, &m.eOff) || up(eN, &m.eN) || up(nodeElem, &m.nodeElem) || up(tOff, &m.tOff) || up(traceOwner, &m.traceOwner) || up(fOff, &m.fOff) ||
            up(faceNodeFace, &m.faceNodeFace) || up(fo, &m.fo) || up(proj, &m.proj) || up(vElemFace, &m.elemFace) || up(vElemFaceSide, &m.elemFaceSide) ||
            up(vFaceElem, &m.faceElem) || up(vFaceElemSide, &m.faceElemSide) || up(vRot, &m.faceRot) || up(vType, &m.faceType) || up(vZone, &m.faceZone) ||
            up(ops, &m.ops) || up(ts, &m.tset)) return 2;
        // geometry: [node][3] x three directions -> Ja [9][nNodes]
        const size_t nn = (size_t)m.nNodes, nf = (size_t)m.nFaceNodes;
        std::vector<double> ja(9 * nn), iJ(nn), tmp;
        for (size_t g = 0; g < nn; ++g) for (int c = 0; c < 3; ++c) {
            ja[(size_t)(0 + c) * nn + g] = jGradXi[3 * g + c]; ja[(size_t)(3 + c) * nn + g] = jGradEta[3 * g + c]; ja[(size_t)(6 + c) * nn + g] = jGradZeta[3 * g + c];
        }
        for (size_t g = 0; g < nn; ++g) iJ[g] = 1.0 / jacobian[g];   // MappedGeometry.f90:382
        std::vector<double> vJ(jacobian, jacobian + nn), vfJ(faceJacobian, faceJacobian + nf);
        if (up(ja, &m.Ja) || up(vJ, &m.J) || up(iJ, &m.invJ) || up(vfJ, &m.fJ)) return 2;
        toSoA(faceNormal, nf, 3, tmp); if (up(tmp, &m.fN)) return 2;
        toSoA(faceT1, nf, 3, tmp); if (up(tmp, &m.fT1)) return 2;
        toSoA(faceT2, nf, 3, tmp); if (up(tmp, &m.fT2)) return 2;
        if (field(&m.Q, 5 * nn) || field(&m.G, 5 * nn) || field(&m.QDot, 5 * nn) || field(&m.Ux, 5 * nn) || field(&m.Uy, 5 * nn) || field(&m.Uz, 5 * nn) ||
            field(&m.Fc, 15 * nn) || field(&m.tr, 15 * (size_t)m.nTrace) || field(&m.fStarE, 5 * (size_t)m.nTrace) || field(&m.unStarE, 15 * (size_t)m.nTrace) ||
            field(&m.fQ, 10 * nf) || field(&m.fU, 30 * nf) || field(&m.fFlux, 15 * nf) || field(&m.partial, 8 * (size_t)std::max(nElem, nFace))) return 2;
        m.partialStride = std::max(nElem, nFace);
        m.S = nullptr; m.dWall = nullptr; m.fDWall = nullptr; m.lesDelta = nullptr; m.fDelta = nullptr;
        if (volume && faceSurface) {
            std::vector<double> dl(nElem), fd(nFace);
            for (int e = 0; e < nElem; ++e) dl[e] = std::pow(volume[e] / (double)(eN[3 * e] * eN[3 * e + 1] * eN[3 * e + 2]), 1.0 / 3.0);
            for (int f = 0; f < nFace; ++f) fd[f] = std::sqrt(faceSurface[f] / (double)((fo[6 * f] + 1) * (fo[6 * f + 1] + 1)));
            if (up(dl, &m.lesDelta) || up(fd, &m.fDelta)) return 2;
        }
        m.volume = nullptr; m.fPenalty = nullptr;
        if (volume) { std::vector<double> v(volume, volume + nElem); if (up(v, &m.volume)) return 2; }
        lesWallModel = physics.les != H3D_LES_NONE && physics.les_wall_model == 1;
        haveMesh = true;
        return 0;
    }
    // f % geom % h (HexMesh.f90:3016-3041) -> the penalty of every face, PenaltyParameterNS with maxval(f % Nf) (EllipticIP.f90:678-687)
    int setFaceH(const H3dPhysics& physics, const double* faceH) {
        if (!haveMesh) return fail("h3d_set_face_h: set the mesh first");
        if (!faceH) return fail("h3d_set_face_h: null array");
        std::vector<double> pen(m.nFace);
        for (int f = 0; f < m.nFace; ++f) {
            const int Nmax = std::max(hFo[6 * (size_t)f], hFo[6 * (size_t)f + 1]);
            pen[f] = 0.5 * physics.penaltyParameter * (Nmax + 1) * (Nmax + 2) / faceH[f];
        }
        return up(pen, &m.fPenalty);
    }
    // e % geom % dWall, f % geom % dWall (HexMesh.f90:5594-5692) in the packed sizes of this mesh
    int setWallDistance(const double* dWallElem, const double* dWallFace) {
        if (!haveMesh) return fail("h3d_set_wall_distance: set the mesh first");
        if (!dWallElem || !dWallFace) return fail("h3d_set_wall_distance: null array");
        std::vector<double> a(dWallElem, dWallElem + m.nNodes), b(dWallFace, dWallFace + m.nFaceNodes);
        if (up(a, &m.dWall) || up(b, &m.fDWall)) return 2;
        return 0;
    }
    int setBoundaryConditions(int nZ, const int* bcType, const double* bcParams) {
        for (int z = 0; z < nZ; ++z) if (bcType[z] < H3D_BC_PERIODIC || bcType[z] > H3D_BC_OUTFLOW) return fail("h3d_set_boundary_conditions: unknown boundary condition type");
        std::vector<int> t(bcType, bcType + nZ); std::vector<double> p(bcParams, bcParams + 16 * (size_t)nZ);
        if (up(t, &m.bcType) || up(p, &m.bcParams)) return 2;
        nZones = nZ; haveBC = true;
        return 0;
    }
    // MPIfaces (MPI_Face.f90:16-57; HexMesh.f90:2571-2668): per neighbour the local MPI faces in exchange order and the local side
    int setHalo(int nNeighbors, const int* neighborRank, const int* faceCount, const int* faceIDs, const int* thisSide) {
        if (!haveMesh) return fail("h3d_set_mesh_p must precede h3d_set_halo");
        const std::vector<long long>& fOff = hFOff; const std::vector<int>& faceType = hFaceType; const std::vector<int>& faceElem = hFaceElem;
        int nHalo = 0;
        for (int b = 0; b < nNeighbors; ++b) { if (faceCount[b] < 1 || neighborRank[b] < 0 || neighborRank[b] >= nranks) return fail("h3d_set_halo: bad neighbour table"); nHalo += faceCount[b]; }
        if (nHalo != nMpiFaces) return fail("h3d_set_halo: the halo must list every MPI face of the mesh exactly once");
        std::vector<int> hFace(faceIDs, faceIDs + nHalo), hSide(thisSide, thisSide + nHalo), hNbr(nHalo), hNodeFace;
        std::vector<long long> hOff(nHalo + 1, 0);
        std::vector<char> seen(m.nFace, 0);
        nbrRank.assign(neighborRank, neighborRank + nNeighbors); nbrNodeOff.assign(nNeighbors + 1, 0);
        int k = 0;
        for (int b = 0; b < nNeighbors; ++b) {
            for (int q = 0; q < faceCount[b]; ++q, ++k) {
                const int f = hFace[k];
                if (f < 0 || f >= m.nFace || faceType[f] != H3D_FACE_MPI || seen[f]) return fail("h3d_set_halo: a halo face is not an MPI face of the mesh (or is listed twice)");
                if (hSide[k] < 0 || hSide[k] > 1 || faceElem[2 * f + hSide[k]] < 0) return fail("h3d_set_halo: thisSide is not the side of the local element");
                seen[f] = 1; hNbr[k] = b;
                hOff[k + 1] = hOff[k] + (fOff[f + 1] - fOff[f]);
                for (long long g = hOff[k]; g < hOff[k + 1]; ++g) hNodeFace.push_back(k);
            }
            nbrNodeOff[b + 1] = hOff[k];
        }
        m.nHalo = nHalo; m.nNbr = nNeighbors; m.nHaloNodes = hOff[nHalo];
        if (up(hFace, &m.haloFace) || up(hSide, &m.haloSide) || up(hNbr, &m.haloNbr) || up(hOff, &m.hOff) || up(hNodeFace, &m.hNodeFace) || up(nbrNodeOff, &m.nbrNodeOff)) return 2;
        if (field(&m.sendBuf, 15 * (size_t)m.nHaloNodes) || field(&m.recvBuf, 15 * (size_t)m.nHaloNodes)) return 2;
        haveHalo = true;
        return 0;
    }
    // the local side's traces of every MPI face to the neighbours, theirs into the remote side (nSets field sets of 5, set stride 10 fields)
    int exchange(int nSets, double* faceField) {
        if (nranks < 2) return 0;
        // every rank calls the backend, also one without neighbours (a backend may synchronise all ranks)
        if (haveHalo) launch(MxHaloPack{m, nSets, faceField, m.sendBuf, 0}, m.nHaloNodes);
        std::vector<long long> off(nbrRank.size()), cnt(nbrRank.size());
        for (size_t b = 0; b < nbrRank.size(); ++b) { off[b] = (long long)nSets * 5 * nbrNodeOff[b]; cnt[b] = (long long)nSets * 5 * (nbrNodeOff[b + 1] - nbrNodeOff[b]); }
        be.exchange(m.sendBuf, m.recvBuf, (int)nbrRank.size(), nbrRank.data(), off.data(), cnt.data());
        if (haveHalo) launch(MxHaloPack{m, nSets, faceField, m.recvBuf, 1}, m.nHaloNodes);
        return check();
    }
    int ready() {
        if (!haveMesh) return fail("no mesh");
        if (nMpiFaces > 0 && !haveHalo) return fail("mesh has MPI faces but h3d_set_halo was not called");
        if (lesWallModel && !m.dWall) return fail("the LES wall model needs the wall distances (h3d_set_wall_distance)");
        if (interiorPenalty && !m.fPenalty) return fail("the interior penalty needs the faces' h (h3d_set_face_h)");
        if (nBoundaryFaces > 0 && !haveBC) return fail("mesh has boundary faces but h3d_set_boundary_conditions was not called");
        if (nBoundaryFaces > 0 && maxZone >= nZones) return fail("a boundary face refers to a zone beyond the table of h3d_set_boundary_conditions");
        return 0;
    }
    // the packed host array crosses PCIe as it is; the transposition runs on the device (staging buffer of one 5-vector field)
    int uploadField(double* dst, const double* src) {
        if (!stage && field(&stage, 5 * (size_t)m.nNodes)) return 2;
        be.upload(stage, src, 5 * (size_t)m.nNodes);
        launch(MxAosToSoa{m, 5, stage, dst}, m.nNodes);
        return check();
    }
    int downloadField(double* dst, const double* src) {
        if (!stage && field(&stage, 5 * (size_t)m.nNodes)) return 2;
        launch(MxSoaToAos{m, 5, src, stage}, m.nNodes);
        be.download(dst, stage, 5 * (size_t)m.nNodes);
        return check();
    }
    int uploadQ(const double* Q) { if (!haveMesh) return fail("no mesh"); return uploadField(m.Q, Q); }
    int download(double* Q, double* QDot, double* Ux, double* Uy, double* Uz) {
        if (!haveMesh) return fail("no mesh");
        double* dst[5] = {Q, QDot, Ux, Uy, Uz}; const double* src[5] = {m.Q, m.QDot, m.Ux, m.Uy, m.Uz};
        for (int a = 0; a < 5; ++a) if (dst[a] && downloadField(dst[a], src[a])) return 2;
        return 0;
    }
    // h3d_snapshot_begin / _end on such meshes: the copy is taken (synchronously) at the point of the time loop where _begin is called
    std::vector<double> snapshot; bool snapPending = false;
    int snapshotBegin() {
        if (!haveMesh) return fail("no mesh");
        if (snapPending) return fail("a snapshot is already in flight: call h3d_snapshot_end first");
        snapshot.resize(5 * (size_t)m.nNodes);
        if (downloadField(snapshot.data(), m.Q)) return 2;
        snapPending = true;
        return 0;
    }
    int snapshotEnd(double* Q) {
        if (!snapPending) return fail("no snapshot in flight");
        snapPending = false;
        if (Q) std::memcpy(Q, snapshot.data(), snapshot.size() * sizeof(double));
        return 0;
    }
    int setSource(const double* S) {
        if (!haveMesh) return fail("no mesh");
        if (!S) { m.S = nullptr; return 0; }
        if (!dSource && field(&dSource, 5 * (size_t)m.nNodes)) return 2;
        if (uploadField(dSource, S)) return 2;
        m.S = dSource;
        return 0;
    }
    // HexMesh_ProlongSolutionToFaces / ProlongGradientsToFaces: traces at the element order, then adaption to the face order
    void prolong(const double* src, double* dstFace) {
        launch(MxTrace{m, 1, {src, nullptr, nullptr}, m.tr}, m.nTrace);
        launch(MxAdapt{m, 1, m.tr, dstFace}, 2 * m.nFaceNodes);
    }
    void prolongGradients() {      // the three gradients in one pass: 15 traces per node
        launch(MxTrace{m, 3, {m.Ux, m.Uy, m.Uz}, m.tr}, m.nTrace);
        launch(MxAdapt{m, 3, m.tr, m.fU}, 2 * m.nFaceNodes);
    }
    // ComputeTimeDerivative (SpatialDiscretization.f90:227-320) followed by the update of one Runge-Kutta stage
    int residual(const H3dPhysics& physics, const MxRk& rk) {
        if (ready()) return 1;
        prolong(m.Q, m.fQ);
        if (exchange(1, m.fQ)) return 2;
        if (physics.computeGradients) {
            launch(MxLocalGrad{m, ph}, m.nNodes);
            // IP_ComputeGradient (EllipticIP.f90:189-362) prolongs the LOCAL gradients, then lifts the interface jumps; BR1 lifts first
            if (interiorPenalty) { prolongGradients(); if (exchange(3, m.fU)) return 2; }
            if (physics.flowIsNavierStokes) {
                launch(MxGradFace{m, ph}, m.nFaceNodes);
                launch(MxProject{m, 15, m.fFlux, m.unStarE, 1.0}, m.nTrace);
                launch(MxLift{m, physics.ipVariant, interiorPenalty ? 1 : 0}, m.nNodes);
            }
            if (!interiorPenalty) { prolongGradients(); if (exchange(3, m.fU)) return 2; }
        }
        launch(MxFlux{m, ph, splitForm ? 1 : 0}, m.nNodes);
        launch(MxRiemann{m, ph}, m.nFaceNodes);
        launch(MxProject{m, 5, m.fFlux, m.fStarE, -1.0}, m.nTrace);
        if (splitForm) {
            launch(MxVolumeSplit{m, ph}, m.nNodes);
            if (rk.mode != 0) launch(MxUpdate{m, rk}, 5 * m.nNodes);
        } else launch(MxVolume{m, rk}, m.nNodes);
        return check();
    }
    int rkStage(const H3dPhysics& physics, int scheme, int k, double dt) {
        MxRk rk;
        if (!mxRkStages(scheme)) return fail("unknown Runge-Kutta scheme");
        if (!mxRkCoefficients(scheme, k, rk, dt)) return fail("Runge-Kutta stage out of range");
        const int rc = residual(physics, rk);
        if (rc || !limited || rk.mode != 2) return rc;
        launch(MxLimiter{m, ph, limiterMin}, m.nElem);      // stage_limiter after every SSPRK stage (ExplicitMethods.f90:1050-1052, 1177-1179)
        return check();
    }
    int enableLimiter(int enabled, double minimum) {
        if (!haveMesh) return fail("h3d_enable_limiter: set the mesh first");
        if (enabled && !m.volume) return fail("the limiter needs the element volumes (h3d_set_mesh_p: volume)");
        limited = enabled != 0;
        if (minimum > 0.0) limiterMin = minimum;
        return 0;
    }
    int statisticsUpdate(const H3dPhysics& physics, int reset) {
        if (ready()) return 1;
        const int nv = physics.computeGradients ? 29 : 14;
        if (!dStats || statVars != nv) { if (field(&dStats, (size_t)nv * m.nNodes)) return 2; statVars = nv; statSamples = 0; reset = 0; }
        if (reset) { be.zero(dStats, (size_t)nv * m.nNodes); statSamples = 0; }
        const double inv = 1.0 / (statSamples + 1), ratio = statSamples * inv;
        launch(MxStatistics{m, nv, ratio, inv, dStats}, m.nNodes);
        ++statSamples;
        return check();
    }
    int statisticsDownload(double* data, int* nVars, int* nSamples) {
        if (!dStats) return fail("no statistics have been accumulated");
        *nVars = statVars; *nSamples = statSamples;
        if (!data) return 0;
        const size_t nn = (size_t)m.nNodes;
        hBuf.resize((size_t)statVars * nn);
        be.download(hBuf.data(), dStats, hBuf.size());
        if (check()) return 2;
        for (size_t g = 0; g < nn; ++g) for (int c = 0; c < statVars; ++c) data[g * statVars + c] = hBuf[(size_t)c * nn + g];
        return 0;
    }
    int rkStep(const H3dPhysics& physics, int scheme, double dt, int ctdAfterStep) {
        const int ns = mxRkStages(scheme);
        if (!ns) return fail("unknown Runge-Kutta scheme");
        for (int k = 0; k < ns; ++k) { const int rc = rkStage(physics, scheme, k, dt); if (rc) return rc; }
        if (ctdAfterStep) return residual(physics, MxRk{0, 0.0, 0.0, 0.0, 0});
        return 0;
    }
    // the first nCols columns of the per-element (per-face) results: hPartial[q * partialStride + e]
    int partials(int nCols) {
        hPartial.resize((size_t)nCols * m.partialStride);
        be.download(hPartial.data(), m.partial, hPartial.size());
        return check();
    }
    double part(int q, long long e) const { return hPartial[(size_t)q * m.partialStride + e]; }
    int maxResiduals(double out[5], int* nanFlag) {
        if (!haveMesh) return fail("no mesh");
        launch(MxRedResidual{m}, m.nElem);
        if (partials(6)) return 2;
        double v[6] = {0, 0, 0, 0, 0, 0};
        for (int e = 0; e < m.nElem; ++e) for (int q = 0; q < 6; ++q) v[q] = std::fmax(v[q], part(q, e));
        if (nranks > 1) { be.allreduce(v, 6, 0); if (check()) return 2; }
        for (int q = 0; q < 5; ++q) out[q] = v[q];
        *nanFlag = v[5] > 0.5 ? 1 : 0;
        return 0;
    }
    int maxTimestep(double cfl, double dcfl, double* dtConv, double* dtVisc) {
        if (ready()) return 1;
        launch(MxRedTimestep{m, ph, cfl, dcfl}, m.nElem);
        if (partials(2)) return 2;
        double a = 1.7976931348623157e308, b = 1.7976931348623157e308;
        for (int e = 0; e < m.nElem; ++e) { a = std::fmin(a, part(0, e)); b = std::fmin(b, part(1, e)); }
        if (nranks > 1) { double ab[2] = {a, b}; be.allreduce(ab, 2, 1); if (check()) return 2; a = ab[0]; b = ab[1]; }
        *dtConv = a; *dtVisc = b;
        return 0;
    }
    int volumeIntegral(const H3dPhysics& physics, int kind, double* val) {
        if (!haveMesh) return fail("no mesh");
        switch (kind) {
            case H3D_INT_VOLUME: case H3D_INT_KINETIC_ENERGY: case H3D_INT_KINETIC_ENERGY_RATE: case H3D_INT_VELOCITY: case H3D_INT_INTERNAL_ENERGY:
            case H3D_INT_ENTROPY: case H3D_INT_MATH_ENTROPY: case H3D_INT_ENTROPY_RATE: break;
            case H3D_INT_ENSTROPHY: case H3D_INT_ENTROPY_BALANCE: case H3D_INT_KINETIC_ENERGY_BALANCE:
                if (!physics.computeGradients) return fail("volume integral needs gradients");
                break;
            default: return fail("unknown volume integral");
        }
        launch(MxRedIntegral{m, ph, kind}, m.nElem);
        if (partials(1)) return 2;
        double v = 0.0;
        for (int e = 0; e < m.nElem; ++e) v = v + part(0, e);
        if (nranks > 1) { be.allreduce(&v, 1, 2); if (check()) return 2; }
        *val = v;
        return 0;
    }
    int surfaceIntegral(const H3dPhysics& physics, int zone, int kind, double out[3]) {
        if (ready()) return 1;
        if (kind < H3D_SURF_SURFACE || kind > H3D_SURF_VISCOUS_FORCE) return fail("unknown surface integral");
        if (zone < 0 || zone >= nZones) return fail("surface integral: zone out of range");
        const bool viscous = kind == H3D_SURF_TOTAL_FORCE || kind == H3D_SURF_VISCOUS_FORCE;
        if (viscous && !physics.computeGradients) return fail("surface integral needs gradients");
        prolong(m.Q, m.fQ);   // the state (and gradients) are prolonged anew, as the reference does (SurfaceIntegrals.f90:57-77)
        if (physics.computeGradients) prolongGradients();
        launch(MxRedSurface{m, ph, zone, kind}, m.nFace);
        if (partials(3)) return 2;
        double v[3] = {0, 0, 0};
        for (int f = 0; f < m.nFace; ++f) for (int d = 0; d < 3; ++d) v[d] = v[d] + part(d, f);
        if (nranks > 1) { be.allreduce(v, 3, 2); if (check()) return 2; }
        for (int d = 0; d < 3; ++d) out[d] = v[d];
        return 0;
    }
    int probe(int nProbes, const int* elem, const int* variable, const double* lxi, const double* leta, const double* lzeta, double* values) {
        if (!haveMesh) return fail("no mesh");
        if (nProbes <= 0) return 0;
        const int ld = maxNodes1D;
        for (int p = 0; p < nProbes; ++p) {
            if (elem[p] < 0 || elem[p] >= m.nElem) return fail("probe element out of range");
            if (variable[p] < H3D_PROBE_PRESSURE || variable[p] > H3D_PROBE_K) return fail("unknown probe variable");
        }
        if (probeCap < (size_t)nProbes) {
            dProbeI = be.template alloc<int>(2 * (size_t)nProbes); dProbeD = be.template alloc<double>((3 * (size_t)ld + 1) * nProbes);
            if (!dProbeI || !dProbeD) return check() ? 2 : fail("device allocation failed");
            probeCap = nProbes;
        }
        std::vector<int> iv(2 * (size_t)nProbes); std::vector<double> L(3 * (size_t)ld * nProbes);
        for (int p = 0; p < nProbes; ++p) { iv[p] = elem[p]; iv[nProbes + p] = variable[p]; }
        std::memcpy(&L[0], lxi, (size_t)ld * nProbes * sizeof(double)); std::memcpy(&L[(size_t)ld * nProbes], leta, (size_t)ld * nProbes * sizeof(double));
        std::memcpy(&L[2 * (size_t)ld * nProbes], lzeta, (size_t)ld * nProbes * sizeof(double));
        be.upload(dProbeI, iv.data(), iv.size()); be.upload(dProbeD, L.data(), L.size());
        double* dVal = dProbeD + 3 * (size_t)ld * nProbes;
        launch(MxProbe{m, ph, dProbeI, dProbeI + nProbes, dProbeD, nProbes, ld, dVal}, nProbes);
        be.download(values, dVal, (size_t)nProbes);
        return check();
    }
};

}  // namespace h3d
