// sm_100a kernels of the explicit NS residual + RK update (FP64, HBM-bound; no tensor-core work: the
// per-element contractions are (N+1)-term sums, far below the FP64 ridge -- see DESIGN.md).
//
// Device layout (all FP64, structure of arrays):
//   element fields  A[c][e][node]   node = (k*n + j)*n + i        -> coalesced per component
//   face fields     A[...][f][m]    m = face-frame node j*n + i   -> coalesced per component
//   fQ    [side][eq][f][m]          prolonged solution, both sides in the FACE frame (FaceClass.f90:281-381)
//   fU    [dir][side][eq][f][m]     prolonged gradients
//   fStar [eq][f][m]                (F*_inv - F*_visc) J_f in the face frame, stored ONCE; the right element
//                                   reads it through the rotation table with a minus sign (FaceClass.f90:681-690)
//   rotmap[r][ab] -> m              element-trace node (a,b) -> face-frame node for rotation index r
//                                   (r = 0 for the left side, the face rotation for the right side)
// Kernels per RK stage: gradient (element) -> riemann (face) -> volume+lift+RK (+ fused prolongation of the
// updated state for the next stage).
#pragma once
#include <cuda_runtime.h>

#include "h3d_physics.cuh"

namespace h3d {

struct DevMesh {
    int n, nElem, nFace;
    // element fields
    double *Q, *G, *QDot, *Ux, *Uy, *Uz;   // [5][nElem][n3]
    const double* S;                       // [5][nElem][n3] or nullptr
    const double* Ja;                      // [9][nElem][n3]: (3*d + c) = c-th Cartesian component of J a^d
    const double *J, *invJ;                // [nElem][n3]
    const double* lesDelta;                // [nElem]  (V/n^3)^(1/3)
    const int* elemFace;                   // [nElem][6] device face id
    const int* elemInfo;                   // [nElem][6] bit0 side | bits1-3 rotation index | bits4-5 face type | bits 8.. zone+1
    // face fields
    double *fQ, *fU, *fStar;
    const double *fN, *fT1, *fT2;          // [3][nFace][n2]
    const double* fJ;                      // [nFace][n2]
    const double* fDelta;                  // [nFace] sqrt(surface/n^2)
    const int* faceInfo;                   // [nFace] bits0-1 type | bits 8.. zone+1 | bit 2: local side for MPI faces
    const int* rotmap;                     // [8][n2]
    // operators (row-major M(i,l)) and their transposes
    const double *hatDT, *DT, *sharpDT;    // transposed: MT[l*n + i] = M(i,l)
    const double *v, *b, *w;               // v,b: [2][n]
    // boundary conditions
    const int* bcType; const double* bcParams;
};

struct RkArgs {
    int mode;        // 0: residual only (QDot stored), 1: RK update G = a G + QDot, Q += cdt G
    int storeQDot;   // also store QDot in mode 1 (last stage: monitors read it)
    int prolong;     // prolong the (updated) Q to the faces at the end
    double a, cdt;
};

__host__ __device__ constexpr int epbFor(int n) { return (n * n * n >= 192) ? 1 : 256 / (n * n * n); }

#define H3D_EIDX(c, e, node) (((size_t)(c) * m.nElem + (e)) * N3 + (node))
#define H3D_FIDX(c, f, mm) (((size_t)(c) * m.nFace + (f)) * N2 + (mm))

// local face helpers: faces 0..5 = FRONT(eta-),BACK(eta+),BOTTOM(zeta-),RIGHT(xi+),TOP(zeta+),LEFT(xi-)
__device__ __forceinline__ int faceAxis(int lf) { return (lf < 2) ? 1 : ((lf == 2 || lf == 4) ? 2 : 0); }
__device__ __forceinline__ int faceEnd(int lf) { return (lf == 1 || lf == 3 || lf == 4) ? 1 : 0; }
// volume node index of trace node (a,b) on local face lf at normal position l
template <int n>
__device__ __forceinline__ int traceNode(int lf, int a, int b, int l) {
    const int ax = faceAxis(lf);
    if (ax == 0) return (b * n + a) * n + l;          // (eta,zeta) face: i = l, j = a, k = b
    if (ax == 1) return (b * n + l) * n + a;          // (xi,zeta) face: i = a, j = l, k = b
    return (l * n + b) * n + a;                       // (xi,eta) face: i = a, j = b, k = l
}

// ---------------------------------------------------------------------------------------------------------
// Prolongation of NV element fields held in shared memory (sF[le][v][node]) to the faces.
// HexElement_ProlongSolutionToFaces / ...GradientsToFaces (HexElementClass.f90:233-372): trace = sum_l A(l) v(l,end),
// accumulated in ascending l from zero; Face_AdaptSolutionToFace: left copies, right is re-indexed.
// ---------------------------------------------------------------------------------------------------------
template <int n, int NV>
__device__ __forceinline__ void prolong_block(const DevMesh& m, const double* sF, const double* sV, double* dst, int e0, int eEnd) {
    constexpr int N2 = n * n, N3 = n * n * n, EPB = epbFor(n);
    const int total = EPB * 6 * NV * N2;
    for (int o = threadIdx.x; o < total; o += blockDim.x) {
        const int ab = o % N2; int r = o / N2;
        const int vv = r % NV; r /= NV;
        const int lf = r % 6; const int le = r / 6;
        const int e = e0 + le;
        if (e >= eEnd) continue;
        const int a = ab % n, b = ab / n;
        const double* src = sF + ((size_t)le * NV + vv) * N3;
        const double* vv_ = sV + faceEnd(lf) * n;
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) acc = acc + src[traceNode<n>(lf, a, b, l)] * vv_[l];
        const int f = m.elemFace[e * 6 + lf];
        const int info = m.elemInfo[e * 6 + lf];
        const int side = info & 1, ridx = (info >> 1) & 7;
        const int mm = m.rotmap[ridx * N2 + ab];
        const int grp = vv / 5, eq = vv % 5;
        dst[H3D_FIDX((grp * 2 + side) * 5 + eq, f, mm)] = acc;
    }
}

// Stand-alone prolongation of Q (first residual after an upload).
template <int n>
__global__ void __launch_bounds__(epbFor(n) * n * n * n) k_prolong_q(DevMesh m, int eBegin, int eEnd) {
    constexpr int N3 = n * n * n, EPB = epbFor(n);
    extern __shared__ double smem[];
    double* sQ = smem;                  // [EPB][5][N3]
    double* sV = sQ + EPB * 5 * N3;     // [2][n]
    const int le = threadIdx.x / N3, node = threadIdx.x % N3;
    const int e0 = eBegin + blockIdx.x * EPB, e = e0 + le;
    if (threadIdx.x < 2 * n) sV[threadIdx.x] = m.v[threadIdx.x];
    if (e < eEnd) {
#pragma unroll
        for (int q = 0; q < 5; ++q) sQ[(le * 5 + q) * N3 + node] = m.Q[H3D_EIDX(q, e, node)];
    }
    __syncthreads();
    prolong_block<n, 5>(m, sQ, sV, m.fQ, e0, eEnd);
}

// ---------------------------------------------------------------------------------------------------------
// BR1 gradient: local gradient + interface lift + prolongation of the gradients.
//   HexElement_ComputeLocalGradient (HexElementClass.f90:427-531)
//   BR1_ComputeElementInterfaceAverage / BR1_ComputeBoundaryFlux (EllipticBR1.f90:571-736), evaluated on the fly
//   per element side from the two prolonged states: u* n J_f = 1/2 (U_R - U_L) J_f n (same value for both sides)
//   BR1_GradientFaceLoop -> VectorWeakIntegrals_StdFace (EllipticBR1.f90:531-569, DGIntegrals.f90:365-443)
//   HexElement_ProlongGradientsToFaces (HexElementClass.f90:304-372)
// ---------------------------------------------------------------------------------------------------------
template <int n>
__global__ void __launch_bounds__(epbFor(n) * n * n * n) k_gradient(DevMesh m, Phys ph, int eBegin, int eEnd) {
    constexpr int N2 = n * n, N3 = n * n * n, EPB = epbFor(n);
    extern __shared__ double smem[];
    double* sA = smem;                       // phase 1: U [EPB][5][N3]; phase 3: grad [EPB][15][N3]
    double* sH = sA + EPB * 15 * N3;         // [EPB][6][5][N2]  uStar*Jf at element-trace nodes
    double* sNrm = sH + EPB * 6 * 5 * N2;    // [EPB][6][4][N2]  face normal (3) and J_f at element-trace nodes
    double* sDT = sNrm + EPB * 6 * 4 * N2;   // [n][n] DT[l*n+i] = D(i,l)
    double* sB = sDT + N2;                   // [2][n]
    double* sV = sB + 2 * n;                 // [2][n]
    const int le = threadIdx.x / N3, node = threadIdx.x % N3;
    const int e0 = eBegin + blockIdx.x * EPB, e = e0 + le;
    const bool active = e < eEnd;
    const int i = node % n, j = (node / n) % n, k = node / N2;
    for (int t = threadIdx.x; t < N2; t += blockDim.x) sDT[t] = m.DT[t];
    if (threadIdx.x < 2 * n) { sB[threadIdx.x] = m.b[threadIdx.x]; sV[threadIdx.x] = m.v[threadIdx.x]; }
    double U[5];
    if (active) {
#pragma unroll
        for (int q = 0; q < 5; ++q) { U[q] = m.Q[H3D_EIDX(q, e, node)]; sA[(le * 5 + q) * N3 + node] = U[q]; }
    }
    // interface data of the six faces at element-trace nodes
    for (int o = threadIdx.x; o < EPB * 6 * N2; o += blockDim.x) {
        const int ab = o % N2; const int lf = (o / N2) % 6; const int l2 = o / (6 * N2);
        const int ee = e0 + l2;
        if (ee >= eEnd) continue;
        const int f = m.elemFace[ee * 6 + lf];
        const int info = m.elemInfo[ee * 6 + lf];
        const int side = info & 1, ridx = (info >> 1) & 7, ftype = (info >> 4) & 3, zone = (info >> 8) - 1;
        const int mm = m.rotmap[ridx * N2 + ab];
        const double Jf = m.fJ[(size_t)f * N2 + mm];
        double nh[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { nh[d] = m.fN[H3D_FIDX(d, f, mm)]; sNrm[((l2 * 6 + lf) * 4 + d) * N2 + ab] = nh[d]; }
        sNrm[((l2 * 6 + lf) * 4 + 3) * N2 + ab] = Jf;
        if (ftype == H3D_FACE_BOUNDARY) {
            // BR1_ComputeBoundaryFlux: unStar = (u* - u_int) n_d J_f ; (u* - u_int) staged, n_d and J_f applied in the lift
            double Qi[5], us[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) Qi[q] = m.fQ[H3D_FIDX(side * 5 + q, f, mm)];
            bc_grad_vars(ph, m.bcType[zone], m.bcParams + 16 * zone, nh, Qi, us);
#pragma unroll
            for (int q = 0; q < 5; ++q) sH[((l2 * 6 + lf) * 5 + q) * N2 + ab] = (us[q] - Qi[q]);
        } else {
            // BR1_ComputeElementInterfaceAverage: uStar = 1/2 (U_R - U_L) J_f ; n_d applied in the lift
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double UL = m.fQ[H3D_FIDX(0 * 5 + q, f, mm)], UR = m.fQ[H3D_FIDX(1 * 5 + q, f, mm)];
                sH[((l2 * 6 + lf) * 5 + q) * N2 + ab] = 0.5 * (UR - UL) * Jf;
            }
        }
    }
    __syncthreads();
    double gx[5], gy[5], gz[5];
    if (active) {
        double Uxi[5] = {0, 0, 0, 0, 0}, Ueta[5] = {0, 0, 0, 0, 0}, Uzeta[5] = {0, 0, 0, 0, 0};
        const double* sU = sA + le * 5 * N3;
#pragma unroll
        for (int l = 0; l < n; ++l) {
            const double dx = sDT[l * n + i], dy = sDT[l * n + j], dz = sDT[l * n + k];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                Uxi[q] = Uxi[q] + sU[q * N3 + (k * n + j) * n + l] * dx;
                Ueta[q] = Ueta[q] + sU[q * N3 + (k * n + l) * n + i] * dy;
                Uzeta[q] = Uzeta[q] + sU[q * N3 + (l * n + j) * n + i] * dz;
            }
        }
        double ja[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) ja[c] = m.Ja[H3D_EIDX(c, e, node)];
        const double iJ = m.invJ[(size_t)e * N3 + node];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            gx[q] = (Uxi[q] * ja[0] + Ueta[q] * ja[3] + Uzeta[q] * ja[6]) * iJ;
            gy[q] = (Uxi[q] * ja[1] + Ueta[q] * ja[4] + Uzeta[q] * ja[7]) * iJ;
            gz[q] = (Uxi[q] * ja[2] + Ueta[q] * ja[5] + Uzeta[q] * ja[8]) * iJ;
        }
        // lift: faceInt_d = sum over faces in the order L,R,FRONT,BACK,BOTTOM,TOP of unStar_d * b
        const int lfOrder[6] = {5, 3, 0, 1, 2, 4};
        const int abOf[6] = {k * n + i, k * n + i, j * n + i, k * n + j, j * n + i, k * n + j};
        const int idxOf[6] = {j, j, k, i, k, i};
        double fx[5], fy[5], fz[5];
#pragma unroll
        for (int s = 0; s < 6; ++s) {
            const int lf = lfOrder[s];
            const int ab = abOf[lf];
            const double bb = sB[faceEnd(lf) * n + idxOf[lf]];
            const double* H = sH + ((le * 6 + lf) * 5) * N2 + ab;
            const double* Nn = sNrm + ((le * 6 + lf) * 4) * N2 + ab;
            const bool bnd = ((m.elemInfo[e * 6 + lf] >> 4) & 3) == H3D_FACE_BOUNDARY;
            const double Jfb = Nn[3 * N2];
            const double n0 = Nn[0], n1 = Nn[N2], n2 = Nn[2 * N2];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double h = H[q * N2];
                double ux, uy, uz;
                if (bnd) { ux = h * n0 * Jfb; uy = h * n1 * Jfb; uz = h * n2 * Jfb; }
                else { ux = h * n0; uy = h * n1; uz = h * n2; }
                if (s == 0) { fx[q] = ux * bb; fy[q] = uy * bb; fz[q] = uz * bb; }
                else { fx[q] = fx[q] + ux * bb; fy[q] = fy[q] + uy * bb; fz[q] = fz[q] + uz * bb; }
            }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            gx[q] = gx[q] + fx[q] * iJ; gy[q] = gy[q] + fy[q] * iJ; gz[q] = gz[q] + fz[q] * iJ;
            m.Ux[H3D_EIDX(q, e, node)] = gx[q]; m.Uy[H3D_EIDX(q, e, node)] = gy[q]; m.Uz[H3D_EIDX(q, e, node)] = gz[q];
        }
    }
    __syncthreads();   // everyone is done reading U from sA
    if (active) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            sA[(le * 15 + q) * N3 + node] = gx[q]; sA[(le * 15 + 5 + q) * N3 + node] = gy[q]; sA[(le * 15 + 10 + q) * N3 + node] = gz[q];
        }
    }
    __syncthreads();
    prolong_block<n, 15>(m, sA, sV, m.fU, e0, eEnd);
}

// ---------------------------------------------------------------------------------------------------------
// Interface fluxes, one thread per face node.
//   computeElementInterfaceFlux / computeMPIFaceFlux / computeBoundaryFlux (SpatialDiscretization.f90:1710-2028)
//   BR1_RiemannSolver (EllipticBR1.f90:816-868), RiemannSolver pointer (RiemannSolvers_NS.f90)
//   compute_viscosity_at_faces (SpatialDiscretization.f90:1345-1398): mu,kappa from the prolonged states
// ---------------------------------------------------------------------------------------------------------
template <int n>
__global__ void __launch_bounds__(128) k_riemann(DevMesh m, Phys ph, int fBegin, int fEnd) {
    constexpr int N2 = n * n;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = fBegin + (int)(t / N2), mm = (int)(t % N2);
    if (f >= fEnd) return;
    const int info = m.faceInfo[f];
    const int ftype = info & 3, zone = (info >> 8) - 1;
    double nh[3], t1[3], t2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { nh[d] = m.fN[H3D_FIDX(d, f, mm)]; t1[d] = 0.0; t2[d] = 0.0; }
    if (ph.riemann != H3D_RIEMANN_ROE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { t1[d] = m.fT1[H3D_FIDX(d, f, mm)]; t2[d] = m.fT2[H3D_FIDX(d, f, mm)]; }
    }
    const double Jf = m.fJ[(size_t)f * N2 + mm];
    double QL[5], QR[5], visc[5] = {0, 0, 0, 0, 0}, inv[5];
    if (ftype == H3D_FACE_BOUNDARY) {
        const int btype = m.bcType[zone]; const double* P = m.bcParams + 16 * zone;
#pragma unroll
        for (int q = 0; q < 5; ++q) { QL[q] = m.fQ[H3D_FIDX(q, f, mm)]; QR[q] = QL[q]; }
        bc_flow_state(ph, btype, P, nh, QR);
        if (ph.ns) {
            double gx[5], gy[5], gz[5], F[5][3], mu, kappa;
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = m.fU[H3D_FIDX((0 * 2 + 0) * 5 + q, f, mm)]; gy[q] = m.fU[H3D_FIDX((1 * 2 + 0) * 5 + q, f, mm)]; gz[q] = m.fU[H3D_FIDX((2 * 2 + 0) * 5 + q, f, mm)]; }
            laminar_mu_kappa(ph, QL, mu, kappa);
            if (ph.les == H3D_LES_SMAGORINSKY) { const double mut = smagorinsky(ph, m.fDelta[f], QL, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux(ph, QL, gx, gy, gz, mu, 0.0, kappa, F);
#pragma unroll
            for (int q = 0; q < 5; ++q) { visc[q] = F[q][0] * nh[0] + F[q][1] * nh[1] + F[q][2] * nh[2]; }
            bc_neumann(btype, P, QL, visc);
        }
        riemann_solver(ph, QL, QR, nh, t1, t2, inv);
    } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) { QL[q] = m.fQ[H3D_FIDX(q, f, mm)]; QR[q] = m.fQ[H3D_FIDX(5 + q, f, mm)]; }
        if (ph.ns) {
            double gx[5], gy[5], gz[5], FL[5][3], FR[5][3], mu, kappa;
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = m.fU[H3D_FIDX((0 * 2 + 0) * 5 + q, f, mm)]; gy[q] = m.fU[H3D_FIDX((1 * 2 + 0) * 5 + q, f, mm)]; gz[q] = m.fU[H3D_FIDX((2 * 2 + 0) * 5 + q, f, mm)]; }
            laminar_mu_kappa(ph, QL, mu, kappa);
            if (ph.les == H3D_LES_SMAGORINSKY) { const double mut = smagorinsky(ph, m.fDelta[f], QL, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux(ph, QL, gx, gy, gz, mu, 0.0, kappa, FL);
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = m.fU[H3D_FIDX((0 * 2 + 1) * 5 + q, f, mm)]; gy[q] = m.fU[H3D_FIDX((1 * 2 + 1) * 5 + q, f, mm)]; gz[q] = m.fU[H3D_FIDX((2 * 2 + 1) * 5 + q, f, mm)]; }
            laminar_mu_kappa(ph, QR, mu, kappa);
            if (ph.les == H3D_LES_SMAGORINSKY) { const double mut = smagorinsky(ph, m.fDelta[f], QR, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux(ph, QR, gx, gy, gz, mu, 0.0, kappa, FR);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double fx = 0.5 * (FL[q][0] + FR[q][0]), fy = 0.5 * (FL[q][1] + FR[q][1]), fz = 0.5 * (FL[q][2] + FR[q][2]);
                visc[q] = fx * nh[0] + fy * nh[1] + fz * nh[2];
            }
        }
        riemann_solver(ph, QL, QR, nh, t1, t2, inv);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) m.fStar[H3D_FIDX(q, f, mm)] = (inv[q] - visc[q]) * Jf;
}

// ---------------------------------------------------------------------------------------------------------
// Volume term + surface lift + 1/J + source + RK update (+ fused prolongation of the updated state).
//   TimeDerivative_VolumetricContribution (SpatialDiscretization.f90:1602-1682)
//   BaseClass_ComputeInnerFluxes (HyperbolicDiscretizationClass.f90:83-152), BR1_ComputeInnerFluxes (EllipticBR1.f90:740-814)
//   ScalarWeakIntegrals_StdVolumeGreen (DGIntegrals.f90:56-87); SplitDG: HyperbolicSplitForm.f90:64-116 + DGIntegrals.f90:92-129
//   TimeDerivative_FacesContribution -> ScalarWeakIntegrals_StdFace (SpatialDiscretization.f90:1686-1701; DGIntegrals.f90:214-273)
//   QDot /= jacobian (:489-491); QDot += S_NS (:632-638); TakeRK3Step element loop (ExplicitMethods.f90:766-771)
// ---------------------------------------------------------------------------------------------------------
template <int n, bool SPLIT>
__global__ void __launch_bounds__(epbFor(n) * n * n * n) k_volume(DevMesh m, Phys ph, RkArgs rk, int eBegin, int eEnd) {
    constexpr int N2 = n * n, N3 = n * n * n, EPB = epbFor(n);
    extern __shared__ double smem[];
    // StandardDG: sF = contravariant total flux [EPB][3][5][N3]
    // SplitDG   : sF = viscous contravariant flux [EPB][3][5][N3]; sQ = state [EPB][5][N3]; sJa = metrics [EPB][9][N3]
    double* sF = smem;                                // SplitDG + Euler: only 5 fields are needed here (prolongation buffer)
    double* sQ = sF + EPB * ((SPLIT && !ph.ns) ? 5 : 15) * N3;
    double* sJa = sQ + (SPLIT ? EPB * 5 * N3 : 0);
    double* sFs = sJa + (SPLIT ? EPB * 9 * N3 : 0);   // [EPB][6][5][N2] fStar at element-trace nodes (signed)
    double* sHatDT = sFs + EPB * 6 * 5 * N2;          // [n][n]
    double* sSharpDT = sHatDT + N2;                   // [n][n]
    double* sB = sSharpDT + N2;                       // [2][n]
    double* sV = sB + 2 * n;                          // [2][n]
    const int le = threadIdx.x / N3, node = threadIdx.x % N3;
    const int e0 = eBegin + blockIdx.x * EPB, e = e0 + le;
    const bool active = e < eEnd;
    const int i = node % n, j = (node / n) % n, k = node / N2;
    for (int t = threadIdx.x; t < N2; t += blockDim.x) { sHatDT[t] = m.hatDT[t]; if (SPLIT) sSharpDT[t] = m.sharpDT[t]; }
    if (threadIdx.x < 2 * n) { sB[threadIdx.x] = m.b[threadIdx.x]; sV[threadIdx.x] = m.v[threadIdx.x]; }
    // interface fluxes of the six faces at element-trace nodes (left: +, right: -, FaceClass.f90:681-690)
    for (int o = threadIdx.x; o < EPB * 6 * 5 * N2; o += blockDim.x) {
        const int ab = o % N2; int r = o / N2;
        const int q = r % 5; r /= 5;
        const int lf = r % 6; const int l2 = r / 6;
        const int ee = e0 + l2;
        if (ee >= eEnd) continue;
        const int f = m.elemFace[ee * 6 + lf];
        const int info = m.elemInfo[ee * 6 + lf];
        const int side = info & 1, ridx = (info >> 1) & 7;
        const double val = m.fStar[H3D_FIDX(q, f, m.rotmap[ridx * N2 + ab])];
        sFs[((l2 * 6 + lf) * 5 + q) * N2 + ab] = side ? -val : val;
    }
    double Q[5], Finv[5][3];
    double ja[9];
    if (active) {
#pragma unroll
        for (int q = 0; q < 5; ++q) Q[q] = m.Q[H3D_EIDX(q, e, node)];
#pragma unroll
        for (int c = 0; c < 9; ++c) ja[c] = m.Ja[H3D_EIDX(c, e, node)];
        double F[5][3];
        euler_flux(ph, Q, F);
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int d = 0; d < 3; ++d) Finv[q][d] = F[q][0] * ja[3 * d + 0] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
        double Fv[5][3];
        if (ph.ns) {
            double gx[5], gy[5], gz[5], mu, kappa;
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = m.Ux[H3D_EIDX(q, e, node)]; gy[q] = m.Uy[H3D_EIDX(q, e, node)]; gz[q] = m.Uz[H3D_EIDX(q, e, node)]; }
            laminar_mu_kappa(ph, Q, mu, kappa);
            if (ph.les == H3D_LES_SMAGORINSKY) { const double mut = smagorinsky(ph, m.lesDelta[e], Q, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux(ph, Q, gx, gy, gz, mu, 0.0, kappa, F);
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int d = 0; d < 3; ++d) Fv[q][d] = F[q][0] * ja[3 * d + 0] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
        } else {
#pragma unroll
            for (int q = 0; q < 5; ++q) { Fv[q][0] = 0.0; Fv[q][1] = 0.0; Fv[q][2] = 0.0; }
        }
        if (!SPLIT) {
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int d = 0; d < 3; ++d) sF[((le * 3 + d) * 5 + q) * N3 + node] = Finv[q][d] - Fv[q][d];
        } else {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                sQ[(le * 5 + q) * N3 + node] = Q[q];
                if (ph.ns) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) sF[((le * 3 + d) * 5 + q) * N3 + node] = Fv[q][d];
                }
            }
#pragma unroll
            for (int c = 0; c < 9; ++c) sJa[(le * 9 + c) * N3 + node] = ja[c];
        }
    }
    __syncthreads();
    double Qn[5];
    if (active) {
        double vol[5] = {0, 0, 0, 0, 0};
        if (!SPLIT) {
            const double* F1 = sF + ((le * 3 + 0) * 5) * N3; const double* F2 = sF + ((le * 3 + 1) * 5) * N3; const double* F3 = sF + ((le * 3 + 2) * 5) * N3;
#pragma unroll
            for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + i];
#pragma unroll
                for (int q = 0; q < 5; ++q) vol[q] = vol[q] + d * F1[q * N3 + (k * n + j) * n + l]; }
#pragma unroll
            for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + j];
#pragma unroll
                for (int q = 0; q < 5; ++q) vol[q] = vol[q] + d * F2[q * N3 + (k * n + l) * n + i]; }
#pragma unroll
            for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + k];
#pragma unroll
                for (int q = 0; q < 5; ++q) vol[q] = vol[q] + d * F3[q * N3 + (l * n + j) * n + i]; }
        } else {
            const double* sQe = sQ + le * 5 * N3; const double* sJe = sJa + le * 9 * N3; const double* sFe = sF + le * 15 * N3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int me = d == 0 ? i : (d == 1 ? j : k);
                const double jaMe[3] = {ja[3 * d], ja[3 * d + 1], ja[3 * d + 2]};
                for (int l = 0; l < n; ++l) {
                    const int other = d == 0 ? (k * n + j) * n + l : (d == 1 ? (k * n + l) * n + i : (l * n + j) * n + i);
                    double fs[5];
                    if (l == me) {
#pragma unroll
                        for (int q = 0; q < 5; ++q) fs[q] = Finv[q][d];
                    } else {
                        double Qo[5], jo[3];
#pragma unroll
                        for (int q = 0; q < 5; ++q) Qo[q] = sQe[q * N3 + other];
#pragma unroll
                        for (int c = 0; c < 3; ++c) jo[c] = sJe[(3 * d + c) * N3 + other];
                        if (l > me) two_point_flux(ph, Q, Qo, jaMe, jo, fs); else two_point_flux(ph, Qo, Q, jo, jaMe, fs);
                    }
                    const double sd = sSharpDT[l * n + me], hd = sHatDT[l * n + me];
                    if (ph.ns) {
#pragma unroll
                        for (int q = 0; q < 5; ++q) vol[q] = vol[q] + sd * fs[q] + hd * sFe[(d * 5 + q) * N3 + other];
                    } else {
#pragma unroll
                        for (int q = 0; q < 5; ++q) vol[q] = vol[q] + sd * fs[q];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) vol[q] = -vol[q];
        }
        // surface integral in the reference's order L,R,FRONT,BACK,BOTTOM,TOP
        const double* Fs = sFs + (le * 6) * 5 * N2;
        const double bL = sB[i], bR = sB[n + i], bF = sB[j], bBk = sB[n + j], bBo = sB[k], bT = sB[n + k];
        const double Jn = m.J[(size_t)e * N3 + node];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            double fi = Fs[(5 * 5 + q) * N2 + k * n + j] * bL;
            fi = fi + Fs[(3 * 5 + q) * N2 + k * n + j] * bR;
            fi = fi + Fs[(0 * 5 + q) * N2 + k * n + i] * bF;
            fi = fi + Fs[(1 * 5 + q) * N2 + k * n + i] * bBk;
            fi = fi + Fs[(2 * 5 + q) * N2 + j * n + i] * bBo;
            fi = fi + Fs[(4 * 5 + q) * N2 + j * n + i] * bT;
            double r = vol[q] - fi;
            r = r / Jn;
            if (m.S) r = r + m.S[H3D_EIDX(q, e, node)];
            if (rk.mode == 0) {
                m.QDot[H3D_EIDX(q, e, node)] = r;
                Qn[q] = Q[q];
            } else {
                if (rk.storeQDot) m.QDot[H3D_EIDX(q, e, node)] = r;
                const double g = rk.a * m.G[H3D_EIDX(q, e, node)] + r;
                m.G[H3D_EIDX(q, e, node)] = g;
                Qn[q] = Q[q] + rk.cdt * g;
                m.Q[H3D_EIDX(q, e, node)] = Qn[q];
            }
        }
    }
    if (rk.prolong) {
        __syncthreads();
        if (active) {
#pragma unroll
            for (int q = 0; q < 5; ++q) sF[(le * 5 + q) * N3 + node] = Qn[q];
        }
        __syncthreads();
        prolong_block<n, 5>(m, sF, sV, m.fQ, e0, eEnd);
    }
}

// shared-memory footprints (bytes)
inline size_t smemProlong(int n) { const int n3 = n * n * n, E = epbFor(n); return sizeof(double) * ((size_t)E * 5 * n3 + 2 * n); }
inline size_t smemGradient(int n) { const int n2 = n * n, n3 = n2 * n, E = epbFor(n); return sizeof(double) * ((size_t)E * 15 * n3 + (size_t)E * 6 * 9 * n2 + n2 + 4 * n); }
inline size_t smemVolume(int n, bool split, bool ns) { const int n2 = n * n, n3 = n2 * n, E = epbFor(n); return sizeof(double) * ((size_t)E * ((split && !ns) ? 5 : 15) * n3 + (split ? (size_t)E * 14 * n3 : 0) + (size_t)E * 30 * n2 + 2 * n2 + 4 * n); }

}  // namespace h3d
