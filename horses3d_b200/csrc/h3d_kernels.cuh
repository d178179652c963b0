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
//
// Element kernels: one node per thread, EPB elements per CTA (one where an element has 128 nodes or more); the work fields the
// contraction reads are unpadded, the buffers the prolongation reads along lines are padded to an odd row length (conflict-free
// strided trace reads).  Every sum keeps the reference's order.  (Two nodes per thread and two CTAs per SM: h3d_kernels2.cuh.)
#pragma once
#include <cuda_runtime.h>

#include "h3d_physics.cuh"
#include "h3d_tma.cuh"
#include "h3d_mma.cuh"

namespace h3d {

struct DevMesh {
    int n, nElem, nFace;
    // element fields
    double *Q, *G, *QDot, *Ux, *Uy, *Uz;   // [5][nElem][n3]
    const double* S;                       // [5][nElem][n3] or nullptr
    const double* Ja;                      // [9][nElem][n3]: (3*d + c) = c-th Cartesian component of J a^d
    const double *J, *invJ;                // [nElem][n3]
    const double* lesDelta;                // [nElem]  (V/n^3)^(1/3)
    const double *dWall, *fDWall;          // [nElem][n3], [nFace][n2] wall distances (LES wall model) or nullptr
    const int* elemFace;                   // [nElem][6] device face id
    const int* elemInfo;                   // [nElem][8] (6 used; 32-byte records for bulk copies) bit0 side | bits1-3 rotation index | bits4-5 face type | bits 8.. zone+1
    const int* elemTrace;                  // [nElem][6][n2] face-field offset f*n2 + rotmap[r][ab] of every element-trace node
    // face fields
    double *fQ, *fU, *fStar;
    const double *fN, *fT1, *fT2;          // [3][nFace][n2]
    const double* fJ;                      // [nFace][n2]
    const double* fDelta;                  // [nFace] sqrt(surface/n^2)
    const double* fH;                      // [nFace] f % geom % h (interior penalty) or nullptr
    const int* faceInfo;                   // [nFace] bits0-1 type | bits 8.. zone+1
    const int* rotmap;                     // [8][n2]
    // operators, transposed: MT[l*n + i] = M(i,l)
    const double *hatDT, *DT, *sharpDT;
    const double *v, *b, *w;               // v,b: [2][n]
    // boundary conditions
    const int* bcType; const double* bcParams;
};

struct RkArgs {
    int mode;        // 0: residual only (QDot stored), 1: low-storage update G = a G + QDot, Q += cdt G (Euler, RK3, RK5, LSERK14-4),
                     // 2: SSP form Q = a G + b Q + cdt QDot with G = Q at the start of the step (SSPRK33, SSPRK43)
    int storeQDot;   // also store QDot in modes 1, 2 (last stage: monitors read it)
    int prolong;     // prolong the (updated) Q to the faces at the end
    double a, cdt;
    double b;        // mode 2
    int copyG;       // mode 2, first stage: G = Q before the update
};

template <int n>
struct KCfg {
    static constexpr int N2 = n * n, N3 = n * n * n;
    static constexpr int NP = (n % 2 == 0) ? n + 1 : n;       // padded row length (odd)
    static constexpr int NS = N2 * NP;                        // padded nodes per shared-memory field
    static constexpr int TPE = N3;                            // threads per element: one node per thread
    static constexpr int EPB = (TPE >= 128) ? 1 : 256 / TPE;  // elements per CTA
    static constexpr int NT = TPE * EPB;                      // threads per CTA
    static constexpr int MINB = (NT <= 256) ? 2 : 1;          // CTAs per SM the register allocation must allow
    // Persistent variant with bulk-async (TMA) prefetch of the next tile's fields: needs 16-byte aligned, 16-byte
    // multiple runs per field (EPB*N3 even) and the staged fields to fit beside the work buffers (n <= 8).
    static constexpr bool TMA_OK = (n % 2 == 0) && (n <= 8);
    __host__ __device__ static constexpr int pidx(int node) { return (node / n) * NP + node % n; }
};

// local face helpers: faces 0..5 = FRONT(eta-),BACK(eta+),BOTTOM(zeta-),RIGHT(xi+),TOP(zeta+),LEFT(xi-)
__device__ __forceinline__ int faceAxis(int lf) { return (lf < 2) ? 1 : ((lf == 2 || lf == 4) ? 2 : 0); }
__device__ __forceinline__ int faceEnd(int lf) { return (lf == 1 || lf == 3 || lf == 4) ? 1 : 0; }

// ---------------------------------------------------------------------------------------------------------
// Prolongation of NV element fields held in shared memory (sF[le][v][padded node]) to the faces.
// HexElement_ProlongSolutionToFaces / ...GradientsToFaces (HexElementClass.f90:233-372): trace = sum_l A(l) v(l,end),
// accumulated in ascending l from zero; Face_AdaptSolutionToFace: left copies, right is re-indexed.
// sTr[le][6][n2]: face-field offset of every element-trace node; sInfo[le][6]: side / type bits of the element's faces.
// ---------------------------------------------------------------------------------------------------------
// Operator entries passed as a kernel parameter: after unrolling, every use is a constant-bank operand of the FP64
// instruction (no shared-memory load, no address arithmetic).  Row-major: M[i*n + l] = M(i,l); v,b: [end*n + l].
template <int n>
struct Ops { double hatD[n * n], D[n * n], v[2 * n], b[2 * n]; };

// traces of field line `src` (n values, stride STRIDE) on the two opposite faces of axis AX
template <int n, int NV, int AX>
__device__ __forceinline__ void prolong_axis(const DevMesh& m, const Ops<n>& ops, const double* __restrict__ sF, const int* __restrict__ sTr,
                                             const int* __restrict__ sInfo, double* __restrict__ dst, int nLocal) {
    using C = KCfg<n>;
    constexpr int N2 = C::N2, NP = C::NP, NS = C::NS, NT = C::NT, EPB = C::EPB, ROWS = NT / N2;
    constexpr int STRIDE = AX == 0 ? 1 : (AX == 1 ? NP : n * NP);
    constexpr int LF0 = AX == 0 ? 5 : (AX == 1 ? 0 : 2), LF1 = AX == 0 ? 3 : (AX == 1 ? 1 : 4);   // LEFT,RIGHT | FRONT,BACK | BOTTOM,TOP
    const size_t fstride = (size_t)m.nFace * N2;
    const int ab = threadIdx.x % N2, a = ab % n, b = ab / n;
    const int base = AX == 0 ? (b * n + a) * NP : (AX == 1 ? (b * n) * NP + a : b * NP + a);
    auto facePtr = [&](int le, int lf) {
        const int side = sInfo[le * 8 + lf] & 1;
        return dst + (size_t)(side * 5) * fstride + sTr[(le * 6 + lf) * N2 + ab];
    };
    auto line = [&](const double* __restrict__ src, double& acc0, double& acc1) {
        acc0 = 0.0; acc1 = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) { const double sv = src[l * STRIDE]; acc0 = acc0 + sv * ops.v[l]; acc1 = acc1 + sv * ops.v[n + l]; }
    };
    if (EPB == 1) {
        double* p0 = facePtr(0, LF0); double* p1 = facePtr(0, LF1);
#pragma unroll 2
        for (int vv = threadIdx.x / N2; vv < NV; vv += ROWS) {
            double acc0, acc1;
            line(sF + vv * NS + base, acc0, acc1);
            const size_t fo = (size_t)((vv / 5) * 10 + vv % 5) * fstride;
            p0[fo] = acc0; p1[fo] = acc1;
        }
    } else {
        for (int row = threadIdx.x / N2; row < nLocal * NV; row += ROWS) {
            const int vv = row % NV, le = row / NV;
            double acc0, acc1;
            line(sF + (le * NV + vv) * NS + base, acc0, acc1);
            const size_t fo = (size_t)((vv / 5) * 10 + vv % 5) * fstride;
            facePtr(le, LF0)[fo] = acc0; facePtr(le, LF1)[fo] = acc1;
        }
    }
}

// All three axes of one field at once (EPB == 1 kernels): six independent accumulation chains per field instead of two, so the
// 8-cycle latency of the dependent DADDs overlaps (the phase was bound by exactly that: short_sb / wait stalls at a third of the
// FP64 and shared-memory rates, profiles/r2_f_bench_default/regions_core.txt).  Per line the summation order is unchanged.
template <int n, int NV>
__device__ __forceinline__ void prolong_fused(const DevMesh& m, const Ops<n>& ops, const double* __restrict__ sF, const int* __restrict__ sTr,
                                              const int* __restrict__ sInfo, double* __restrict__ dst) {
    using C = KCfg<n>;
    constexpr int N2 = C::N2, NP = C::NP, NS = C::NS, NT = C::NT, ROWS = NT / N2;
    const size_t fstride = (size_t)m.nFace * N2;
    const int ab = threadIdx.x % N2, a = ab % n, b = ab / n;
    const int bx = (b * n + a) * NP, by = (b * n) * NP + a, bz = b * NP + a;
    const int lfs[6] = {5, 3, 0, 1, 2, 4};   // faces LEFT, RIGHT | FRONT, BACK | BOTTOM, TOP
    int off[6];                              // trace offset, side in the sign bit (kept as six ints: six pointers spill)
#pragma unroll
    for (int s = 0; s < 6; ++s) off[s] = sTr[lfs[s] * N2 + ab] | ((sInfo[lfs[s]] & 1) << 31);
#pragma unroll 1
    for (int vv = threadIdx.x / N2; vv < NV; vv += ROWS) {
        const double* src = sF + vv * NS;
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int l = 0; l < n; ++l) {
            const double sx = src[bx + l], sy = src[by + l * NP], sz = src[bz + l * n * NP];
            const double v0 = ops.v[l], v1 = ops.v[n + l];
            acc[0] = acc[0] + sx * v0; acc[1] = acc[1] + sx * v1;
            acc[2] = acc[2] + sy * v0; acc[3] = acc[3] + sy * v1;
            acc[4] = acc[4] + sz * v0; acc[5] = acc[5] + sz * v1;
        }
        const size_t fo = (size_t)((vv / 5) * 10 + vv % 5) * fstride;
#pragma unroll
        for (int s = 0; s < 6; ++s) dst[fo + (off[s] < 0 ? 5 * fstride : 0) + (size_t)(off[s] & 0x7fffffff)] = acc[s];
    }
}

template <int n, int NV>
__device__ __forceinline__ void prolong_block(const DevMesh& m, const Ops<n>& ops, const double* __restrict__ sF,
                                              const int* __restrict__ sTr, const int* __restrict__ sInfo, double* __restrict__ dst, int nLocal) {
    // One work item = (element, axis, field, trace node): the line of n values along the axis is read once from shared
    // memory and contracted with both end vectors, giving the traces on the two opposite faces of that axis.
    // The trace node ab of a thread is fixed (NT is a multiple of n^2); items advance over (element, field).
    static_assert(KCfg<n>::NT % KCfg<n>::N2 == 0, "threads per CTA must be a multiple of n^2");
    if (KCfg<n>::EPB == 1) { prolong_fused<n, NV>(m, ops, sF, sTr, sInfo, dst); return; }
    prolong_axis<n, NV, 0>(m, ops, sF, sTr, sInfo, dst, nLocal);
    prolong_axis<n, NV, 1>(m, ops, sF, sTr, sInfo, dst, nLocal);
    prolong_axis<n, NV, 2>(m, ops, sF, sTr, sInfo, dst, nLocal);
}

// BR2: prolongation of the 15 LOCAL gradient fields followed by the interface-gradient correction of the element's own
// side (BR2_ComputeGradientFaceIntegrals, EllipticBR2.f90:356-451): trace -= eta * unStar_d * (b v)(l) / J(l), l ascending.
// The reference indexes the face storage with the ELEMENT's trace indices, so the face-frame node m receives the correction
// evaluated at the element-trace node of the same index m (identical for unrotated faces); restated as is.
// sH [le][6][5][N2] = Uhat, sNrm [le][6][4][N2] = normal, J_f at element-trace nodes.
template <int n, int AX>
__device__ __forceinline__ void prolong_axis_br2(const DevMesh& m, const Phys& ph, const Ops<n>& ops, const double* __restrict__ sF, const int* __restrict__ sTr,
                                                 const int* __restrict__ sInfo, const double* __restrict__ sH, const double* __restrict__ sNrm,
                                                 double* __restrict__ dst, int e0, int nLocal) {
    using C = KCfg<n>;
    constexpr int N2 = C::N2, N3 = C::N3, NP = C::NP, NS = C::NS, NT = C::NT, ROWS = NT / N2, NV = 15;
    constexpr int STRIDE = AX == 0 ? 1 : (AX == 1 ? NP : n * NP);
    constexpr int GSTRIDE = AX == 0 ? 1 : (AX == 1 ? n : N2);
    constexpr int LF0 = AX == 0 ? 5 : (AX == 1 ? 0 : 2), LF1 = AX == 0 ? 3 : (AX == 1 ? 1 : 4);
    const size_t fstride = (size_t)m.nFace * N2;
    const int ab = threadIdx.x % N2, a = ab % n, b = ab / n;
    const int base = AX == 0 ? (b * n + a) * NP : (AX == 1 ? (b * n) * NP + a : b * NP + a);
    for (int row = threadIdx.x / N2; row < nLocal * NV; row += ROWS) {
        const int vv = row % NV, le = row / NV, d = vv / 5, q = vv % 5;
        const double* src = sF + (le * NV + vv) * NS + base;
        double acc[2] = {0.0, 0.0};
#pragma unroll
        for (int l = 0; l < n; ++l) { const double sv = src[l * STRIDE]; acc[0] = acc[0] + sv * ops.v[l]; acc[1] = acc[1] + sv * ops.v[n + l]; }
#pragma unroll
        for (int end = 0; end < 2; ++end) {
            const int lf = end ? LF1 : LF0;
            const int off = sTr[(le * 6 + lf) * N2 + ab];
            const int mm = off % N2, am = mm % n, bm = mm / n;
            const double un = sH[((le * 6 + lf) * 5 + q) * N2 + mm] * sNrm[((le * 6 + lf) * 4 + d) * N2 + mm];
            const double* iJ = m.invJ + (size_t)(e0 + le) * N3 + (AX == 0 ? (bm * n + am) * n : (AX == 1 ? (bm * n) * n + am : bm * n + am));
            const double eu = ph.eta * un;
            double t = acc[end];
#pragma unroll
            for (int l = 0; l < n; ++l) t = t - eu * (ops.b[end * n + l] * ops.v[end * n + l]) * iJ[l * GSTRIDE];
            const int side = sInfo[le * 8 + lf] & 1;
            dst[(size_t)(side * 5) * fstride + off + (size_t)((vv / 5) * 10 + vv % 5) * fstride] = t;
        }
    }
}

template <int n>
__device__ __forceinline__ void load_face_tables(const DevMesh& m, int* sTr, int* sInfo, int e0, int nLocal) {
    for (int t = threadIdx.x; t < nLocal * 8; t += blockDim.x) sInfo[t] = m.elemInfo[(size_t)e0 * 8 + t];
    for (int t = threadIdx.x; t < nLocal * 6 * n * n; t += blockDim.x) sTr[t] = m.elemTrace[(size_t)e0 * 6 * n * n + t];
}

// Stand-alone prolongation of Q (first residual after an upload).
template <int n>
__global__ void __launch_bounds__(KCfg<n>::NT) k_prolong_q(DevMesh m, const __grid_constant__ Ops<n> ops, int eBegin, int eEnd) {
    using C = KCfg<n>;
    constexpr int N3 = C::N3, NS = C::NS, EPB = C::EPB, TPE = C::TPE;
    extern __shared__ double smem[];
    double* sQ = smem;                  // [EPB][5][NS]
    double* sV = sQ + EPB * 5 * NS;     // [2][n]
    int* sTr = (int*)(sV + 2 * n);      // [EPB][6][N2]
    int* sInfo = sTr + EPB * 6 * C::N2; // [EPB][8]
    const int le = threadIdx.x / TPE, tn = threadIdx.x % TPE;
    const int e0 = eBegin + blockIdx.x * EPB, e = e0 + le;
    const int nLocal = min(EPB, eEnd - e0);
    if (threadIdx.x < 2 * n) sV[threadIdx.x] = m.v[threadIdx.x];
    load_face_tables<n>(m, sTr, sInfo, e0, nLocal);
    const size_t es = (size_t)m.nElem * N3;
    if (e < eEnd) {
        {
            const int node = tn;
            {
                const double* q = m.Q + (size_t)e * N3 + node;
#pragma unroll
                for (int c = 0; c < 5; ++c) sQ[(le * 5 + c) * NS + C::pidx(node)] = q[c * es];
            }
        }
    }
    __syncthreads();
    prolong_block<n, 5>(m, ops, sQ, sTr, sInfo, m.fQ, nLocal);
}

// ---------------------------------------------------------------------------------------------------------
// BR1 gradient: local gradient + interface lift + prolongation of the gradients.
//   HexElement_ComputeLocalGradient (HexElementClass.f90:427-531)
//   BR1_ComputeElementInterfaceAverage / BR1_ComputeBoundaryFlux (EllipticBR1.f90:571-736), evaluated on the fly
//   per element side from the two prolonged states: u* n J_f = 1/2 (U_R - U_L) J_f n (same value for both sides)
//   BR1_GradientFaceLoop -> VectorWeakIntegrals_StdFace (EllipticBR1.f90:531-569, DGIntegrals.f90:365-443)
//   HexElement_ProlongGradientsToFaces (HexElementClass.f90:304-372)
// Shared memory: phase 1 {U [5][NS], uStar [6][5][N2], normal+J_f [6][4][N2]}; phase 2 reuses all of it for the
// 15 gradient fields that are prolonged to the faces.
// ---------------------------------------------------------------------------------------------------------
template <int n, bool TMA, bool VISC = false>
struct GradSmem {
    using C = KCfg<n>;
    static_assert(!(TMA && VISC), "the BR2 / IP variant uses plain loads");
    // non-TMA: phase 1 {U [5][NS], interface [6][9][N2]} aliased by phase 2 {grad [15][NS]}
    // TMA    : staged inputs {Q 5, Ja 9, 1/J 1} [15][EPB*N3] + interface [6][9][N2] + grad [15][NS] (no aliasing)
    static constexpr int phase1 = C::EPB * (5 * C::NS + 6 * 9 * C::N2);
    static constexpr int phase2 = C::EPB * 15 * C::NS;
    // BR2 / IP: the interface data stays live until the prolongation, no aliasing
    static constexpr int fields = TMA ? C::EPB * (15 * C::N3 + 6 * 9 * C::N2 + 15 * C::NS) : (VISC ? phase1 + phase2 : (phase1 > phase2 ? phase1 : phase2));
    static constexpr size_t bytes = sizeof(double) * (fields + C::N2 + 4 * n) + sizeof(int) * (TMA ? 2 : 1) * C::EPB * (6 * C::N2 + 8) + 32;
};

// interface data of one element-trace node: raw loads (issued early) and their reduction to uStar / normal / J_f
struct GradIface { double QL[5], QR[5], nh[3], Jf; int info; };

template <int n>
__device__ __forceinline__ void grad_iface_load(const DevMesh& m, int e, int lf, int ab, GradIface& g) {
    constexpr int N2 = n * n;
    const size_t fs = (size_t)m.nFace * N2;
    g.info = m.elemInfo[(size_t)e * 8 + lf];
    const size_t fo = (size_t)m.elemTrace[((size_t)e * 6 + lf) * N2 + ab];
    g.Jf = m.fJ[fo];
#pragma unroll
    for (int d = 0; d < 3; ++d) g.nh[d] = m.fN[d * fs + fo];
#pragma unroll
    for (int q = 0; q < 5; ++q) { g.QL[q] = m.fQ[(size_t)q * fs + fo]; g.QR[q] = m.fQ[(size_t)(5 + q) * fs + fo]; }
}
template <int n, bool VISC = false>
__device__ __forceinline__ void grad_iface_store(const DevMesh& m, const Phys& ph, const GradIface& g, double* hh, double* nrm) {
    constexpr int N2 = n * n;
    const int side = g.info & 1, ftype = (g.info >> 4) & 3, zone = (g.info >> 8) - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) nrm[d * N2] = g.nh[d];
    nrm[3 * N2] = g.Jf;
    if (VISC) {   // the general variant: BR1 / BR2 / IP with any gradient variables
        double Qi[5], Qe[5], UL[5], UR[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) { Qi[q] = g.QL[q]; Qe[q] = g.QR[q]; }
        if (ftype == H3D_FACE_BOUNDARY) {
#pragma unroll
            for (int q = 0; q < 5; ++q) { Qi[q] = side ? g.QR[q] : g.QL[q]; Qe[q] = Qi[q]; }
        }
        get_gradients(ph, Qi, UL);
        if (ph.viscous == H3D_VISCOUS_BR1) {
            if (ftype == H3D_FACE_BOUNDARY) {   // BR1_ComputeBoundaryFlux (EllipticBR1.f90:686-736)
#pragma unroll
                for (int q = 0; q < 5; ++q) UR[q] = UL[q];
                bc_grad_vars<true>(ph, m.bcType[zone], m.bcParams + 16 * zone, g.nh, Qi, UR);
#pragma unroll
                for (int q = 0; q < 5; ++q) hh[q * N2] = (UR[q] - UL[q]);
            } else {                            // BR1_ComputeElementInterfaceAverage (:571-627)
                get_gradients(ph, Qe, UR);
#pragma unroll
                for (int q = 0; q < 5; ++q) hh[q * N2] = 0.5 * (UR[q] - UL[q]) * g.Jf;
            }
        } else {
            // BR2_/IP_GradientInterfaceSolution[Boundary] (EllipticBR2.f90:458-592, EllipticIP.f90:410-585):
            // Uhat = 1/2 (UL - UR) J_f, the boundary state from StateForEqn (FlowState); n_d applied in the lift
            if (ftype == H3D_FACE_BOUNDARY) bc_flow_state(ph, m.bcType[zone], m.bcParams + 16 * zone, g.nh, Qe);
            get_gradients(ph, Qe, UR);
#pragma unroll
            for (int q = 0; q < 5; ++q) hh[q * N2] = 0.5 * (UL[q] - UR[q]) * g.Jf;
        }
        return;
    }
    if (ftype == H3D_FACE_BOUNDARY) {
        // BR1_ComputeBoundaryFlux: unStar = (u* - u_int) n_d J_f ; (u* - u_int) staged, n_d and J_f applied in the lift
        double Qi[5], us[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) { Qi[q] = side ? g.QR[q] : g.QL[q]; us[q] = Qi[q]; }
        bc_grad_vars(ph, m.bcType[zone], m.bcParams + 16 * zone, g.nh, Qi, us);
#pragma unroll
        for (int q = 0; q < 5; ++q) hh[q * N2] = (us[q] - Qi[q]);
    } else {
        // BR1_ComputeElementInterfaceAverage: uStar = 1/2 (U_R - U_L) J_f ; n_d applied in the lift
#pragma unroll
        for (int q = 0; q < 5; ++q) hh[q * N2] = 0.5 * (g.QR[q] - g.QL[q]) * g.Jf;
    }
}

template <int n, bool TMA, bool VISC = false, bool MMA = false>
__global__ void __launch_bounds__(KCfg<n>::NT, KCfg<n>::MINB) k_gradient(DevMesh m, Phys ph, const __grid_constant__ Ops<n> ops, int eBegin, int eEnd) {
    using C = KCfg<n>;
    static_assert(!MMA || (n == 8 && TMA && !VISC && C::EPB == 1), "the DMMA contraction is written for the staged n = 8 BR1 kernel");
    constexpr int N2 = C::N2, N3 = C::N3, NP = C::NP, NS = C::NS, EPB = C::EPB, TPE = C::TPE, NT = C::NT;
    constexpr int TN3 = EPB * N3;                  // nodes of one tile
    constexpr int UW = TMA ? N3 : NS;              // per-field stride of the state in shared memory
    constexpr int ULD = TMA ? n : NP;              // row length of the state in shared memory
    constexpr int IFI = (EPB * 6 * N2 + NT - 1) / NT;   // interface items per thread
    extern __shared__ __align__(16) double smem[];
    double* sIn = smem;                                         // TMA: [15][TN3] = Q 5, Ja 9, 1/J
    double* sU = TMA ? sIn : smem;                              // state: TMA [5][EPB][N3] (field-major), else [EPB][5][NS]
    double* sH = smem + (TMA ? 15 * TN3 : EPB * 5 * NS);        // [EPB][6][5][N2]
    double* sNrm = sH + EPB * 6 * 5 * N2;                       // [EPB][6][4][N2]
    double* sG = (TMA || VISC) ? sNrm + EPB * 6 * 4 * N2 : smem;   // [EPB][15][NS]
    double* sDT = smem + GradSmem<n, TMA, VISC>::fields;        // [n][n]
    double* sB = sDT + N2;
    double* sV = sB + 2 * n;
    constexpr int TABI = EPB * (6 * N2 + 8);                    // ints of one face-table set: trace offsets [EPB][6][N2] + info [EPB][8]
    int* sTab = (int*)(sV + 2 * n);                             // TMA: two sets, the next tile's is prefetched by bulk copies
    uint64_t* bar = (uint64_t*)(((uintptr_t)(sTab + (TMA ? 2 : 1) * TABI) + 7) & ~(uintptr_t)7);   // bar[0]: fields, bar[1..2]: face tables
    const int le = threadIdx.x / TPE, tn = threadIdx.x % TPE;
    const size_t es = (size_t)m.nElem * N3;
    const int nTiles = (eEnd - eBegin + EPB - 1) / EPB;
    for (int t = threadIdx.x; t < N2; t += blockDim.x) sDT[t] = m.DT[t];
    if (threadIdx.x < 2 * n) { sB[threadIdx.x] = m.b[threadIdx.x]; sV[threadIdx.x] = m.v[threadIdx.x]; }
    auto issueTab = [&](int tile, int buf) {   // thread 0: face tables of the tile
        const int e0 = eBegin + tile * EPB;
        const int nLoc = min(EPB, eEnd - e0);
        int* dst = sTab + buf * TABI;
        mbar_arrive_expect_tx(bar + 1 + buf, (uint32_t)(nLoc * (6 * N2 + 8) * sizeof(int)));
        bulk_g2s(dst, m.elemTrace + (size_t)e0 * 6 * N2, (uint32_t)(nLoc * 6 * N2 * sizeof(int)), bar + 1 + buf);
        bulk_g2s(dst + EPB * 6 * N2, m.elemInfo + (size_t)e0 * 8, (uint32_t)(nLoc * 8 * sizeof(int)), bar + 1 + buf);
    };
    // bulk copies of the tile's 15 fields: lane 0 of warp w issues copy w (one thread issuing all of them would keep its
    // warp behind the others at the next barrier); the transaction count is posted by thread 0 alone
    auto issue = [&](int tile) {
        if (threadIdx.x & 31) return;
        const int e0 = eBegin + tile * EPB;
        const int nLoc = min(EPB, eEnd - e0);
        const uint32_t bytes = (uint32_t)(nLoc * N3 * sizeof(double));
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, 15 * bytes);
        for (int c = threadIdx.x >> 5; c < 15; c += (NT + 31) / 32) {
            const double* src = c < 5 ? m.Q + c * es : (c < 14 ? m.Ja + (c - 5) * es : m.invJ);
            bulk_g2s(sIn + c * TN3, src + (size_t)e0 * N3, bytes, bar);
        }
    };
    if (TMA) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); fence_barrier_init(); }
        __syncthreads();
        if ((int)blockIdx.x < nTiles) { if (threadIdx.x == 0) issueTab(blockIdx.x, 0); issue(blockIdx.x); }
    }
    uint32_t parity = 0;
    GradIface gi[IFI];
    bool havePrefetch = false;
    int iter = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++iter) {
        const int e0 = eBegin + tile * EPB, e = e0 + le;
        const int nLocal = min(EPB, eEnd - e0);
        const bool active = e < eEnd;
        const int tbuf = TMA ? (iter & 1) : 0;
        int* sTr = sTab + tbuf * TABI;                          // [EPB][6][N2]
        int* sInfo = sTr + EPB * 6 * N2;                        // [EPB][8]
        if (TMA) {   // as in k_volume: the other table set was last read by the previous tile's prolongation
            if (threadIdx.x == 0 && tile + (int)gridDim.x < nTiles) issueTab(tile + gridDim.x, tbuf ^ 1);
        } else {
            load_face_tables<n>(m, sTr, sInfo, e0, nLocal);
        }
        // interface data of the six faces at element-trace nodes (prefetched during the previous tile when possible)
        if (!havePrefetch) {
#pragma unroll
            for (int it = 0; it < IFI; ++it) {
                const int o = threadIdx.x + it * NT;
                if (o < nLocal * 6 * N2) grad_iface_load<n>(m, e0 + o / (6 * N2), (o / N2) % 6, o % N2, gi[it]);
            }
        }
#pragma unroll
        for (int it = 0; it < IFI; ++it) {
            const int o = threadIdx.x + it * NT;
            if (o < nLocal * 6 * N2) {
                const int ab = o % N2, lf = (o / N2) % 6, l2 = o / (6 * N2);
                grad_iface_store<n, VISC>(m, ph, gi[it], sH + ((l2 * 6 + lf) * 5) * N2 + ab, sNrm + ((l2 * 6 + lf) * 4) * N2 + ab);
            }
        }
        if (!TMA) {
            if (active) {
                {
                    const int node = tn;
                    {
                        const double* q = m.Q + (size_t)e * N3 + node;
                        if (VISC) {   // GetGradientValues at every node (HexElementClass.f90:466-470)
                            double Qn[5], Un[5];
#pragma unroll
                            for (int c = 0; c < 5; ++c) Qn[c] = q[c * es];
                            get_gradients(ph, Qn, Un);
#pragma unroll
                            for (int c = 0; c < 5; ++c) sU[(le * 5 + c) * NS + C::pidx(node)] = Un[c];
                        } else {
#pragma unroll
                            for (int c = 0; c < 5; ++c) sU[(le * 5 + c) * NS + C::pidx(node)] = q[c * es];
                        }
                    }
                }
            }
        } else {
            mbar_wait(bar + 1 + tbuf, (uint32_t)((iter >> 1) & 1));
            mbar_wait(bar, parity); parity ^= 1;
        }
        __syncthreads();
        if constexpr (MMA) mma_gradient_contract<NT / 32, TN3, NS>(sU, sG, sDT);   // U_xi, U_eta, U_zeta of the 5 variables -> sG
        double g[15];
        if (active) {
            {
                const int node = tn;
                {
                    const int i = node % n, j = (node / n) % n, k = node / N2;
                    double Uxi[5] = {0, 0, 0, 0, 0}, Ueta[5] = {0, 0, 0, 0, 0}, Uzeta[5] = {0, 0, 0, 0, 0};
                    // state field q of this element: TMA [q][le][node], else [le][q][padded node]
                    const double* sUe = TMA ? sU + le * N3 : sU + le * 5 * NS;
                    constexpr int QS = TMA ? TN3 : NS;
                    const int bx = (k * n + j) * ULD, by = (k * n) * ULD + i, bz = j * ULD + i;
                    if (MMA) {
                        const int p = C::pidx(node);
#pragma unroll
                        for (int q = 0; q < 5; ++q) { Uxi[q] = sG[q * NS + p]; Ueta[q] = sG[(5 + q) * NS + p]; Uzeta[q] = sG[(10 + q) * NS + p]; }
                    } else {
#pragma unroll
                        for (int l = 0; l < n; ++l) {
                            const double dx = sDT[l * n + i], dy = sDT[l * n + j], dz = sDT[l * n + k];
#pragma unroll
                            for (int q = 0; q < 5; ++q) {
                                Uxi[q] = Uxi[q] + sUe[q * QS + bx + l] * dx;
                                Ueta[q] = Ueta[q] + sUe[q * QS + by + l * ULD] * dy;
                                Uzeta[q] = Uzeta[q] + sUe[q * QS + bz + l * n * ULD] * dz;
                            }
                        }
                    }
                    double ja[9], iJ;
                    if (TMA) {
#pragma unroll
                        for (int c = 0; c < 9; ++c) ja[c] = sIn[(5 + c) * TN3 + le * N3 + node];
                        iJ = sIn[14 * TN3 + le * N3 + node];
                    } else {
                        const double* jap = m.Ja + (size_t)e * N3 + node;
#pragma unroll
                        for (int c = 0; c < 9; ++c) ja[c] = jap[c * es];
                        iJ = m.invJ[(size_t)e * N3 + node];
                    }
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        g[q] = (Uxi[q] * ja[0] + Ueta[q] * ja[3] + Uzeta[q] * ja[6]) * iJ;
                        g[5 + q] = (Uxi[q] * ja[1] + Ueta[q] * ja[4] + Uzeta[q] * ja[7]) * iJ;
                        g[10 + q] = (Uxi[q] * ja[2] + Ueta[q] * ja[5] + Uzeta[q] * ja[8]) * iJ;
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
                        const bool bnd = (!VISC || ph.viscous == H3D_VISCOUS_BR1) && ((sInfo[le * 8 + lf] >> 4) & 3) == H3D_FACE_BOUNDARY;
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
                    double* ox = m.Ux + (size_t)e * N3 + node; double* oy = m.Uy + (size_t)e * N3 + node; double* oz = m.Uz + (size_t)e * N3 + node;
                    if (VISC) {
                        // BR2 / IP prolong the LOCAL gradients (EllipticBR2.f90:160-190, EllipticIP.f90:215-262); the elements get
                        // U -= faceInt * (1/J) (BR2, :332-338) or U += faceInt * (IPmethod * invJacobian) (IP, :396-402)
                        const int p = C::pidx(node);
                        if (ph.viscous != H3D_VISCOUS_BR1) {
#pragma unroll
                            for (int c = 0; c < 15; ++c) sG[(le * 15 + c) * NS + p] = g[c];
                        }
                        const double iJs = ph.ipVariant * iJ;
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            if (ph.viscous == H3D_VISCOUS_BR1) { g[q] = g[q] + fx[q] * iJ; g[5 + q] = g[5 + q] + fy[q] * iJ; g[10 + q] = g[10 + q] + fz[q] * iJ; }
                            else if (ph.viscous == H3D_VISCOUS_BR2) { g[q] = g[q] - fx[q] * iJ; g[5 + q] = g[5 + q] - fy[q] * iJ; g[10 + q] = g[10 + q] - fz[q] * iJ; }
                            else { g[q] = g[q] + fx[q] * iJs; g[5 + q] = g[5 + q] + fy[q] * iJs; g[10 + q] = g[10 + q] + fz[q] * iJs; }
                            ox[q * es] = g[q]; oy[q * es] = g[5 + q]; oz[q * es] = g[10 + q];
                        }
                        if (ph.viscous == H3D_VISCOUS_BR1) {   // BR1 prolongs the lifted gradients
#pragma unroll
                            for (int c = 0; c < 15; ++c) sG[(le * 15 + c) * NS + p] = g[c];
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            // Euler with "compute gradients": local gradient only (base-class ComputeGradient, EllipticDiscretizationClass.f90:122-187)
                            if (ph.ns) { g[q] = g[q] + fx[q] * iJ; g[5 + q] = g[5 + q] + fy[q] * iJ; g[10 + q] = g[10 + q] + fz[q] * iJ; }
                            ox[q * es] = g[q]; oy[q * es] = g[5 + q]; oz[q * es] = g[10 + q];
                        }
                    }
                    if (TMA) {   // the gradient buffer does not alias the inputs: store right away
                        const int p = C::pidx(node);
#pragma unroll
                        for (int c = 0; c < 15; ++c) sG[(le * 15 + c) * NS + p] = g[c];
                    }
                }
            }
        }
        __syncthreads();   // U and the interface data are consumed
        const int next = tile + gridDim.x;
        if (TMA) {
            if (next < nTiles) issue(next);
        } else if (!VISC) {
            if (active) {
                {
                    const int node = tn;
                    {
                        const int p = C::pidx(node);
#pragma unroll
                        for (int c = 0; c < 15; ++c) sG[(le * 15 + c) * NS + p] = g[c];
                    }
                }
            }
            __syncthreads();
        }
        // start the gather of the next tile's interface data; it completes behind the prolongation below
        havePrefetch = false;
        if (next < nTiles) {
            const int e0n = eBegin + next * EPB, nLocN = min(EPB, eEnd - e0n);
#pragma unroll
            for (int it = 0; it < IFI; ++it) {
                const int o = threadIdx.x + it * NT;
                if (o < nLocN * 6 * N2) grad_iface_load<n>(m, e0n + o / (6 * N2), (o / N2) % 6, o % N2, gi[it]);
            }
            havePrefetch = true;
        }
        if (VISC && ph.viscous == H3D_VISCOUS_BR2) {
            prolong_axis_br2<n, 0>(m, ph, ops, sG, sTr, sInfo, sH, sNrm, m.fU, e0, nLocal);
            prolong_axis_br2<n, 1>(m, ph, ops, sG, sTr, sInfo, sH, sNrm, m.fU, e0, nLocal);
            prolong_axis_br2<n, 2>(m, ph, ops, sG, sTr, sInfo, sH, sNrm, m.fU, e0, nLocal);
        } else {
            prolong_block<n, 15>(m, ops, sG, sTr, sInfo, m.fU, nLocal);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Interface fluxes, one thread per face node.
//   computeElementInterfaceFlux / computeMPIFaceFlux / computeBoundaryFlux (SpatialDiscretization.f90:1710-2028)
//   BR1_RiemannSolver (EllipticBR1.f90:816-868), RiemannSolver pointer (RiemannSolvers_NS.f90)
//   compute_viscosity_at_faces (SpatialDiscretization.f90:1345-1398): mu,kappa from the prolonged states
// ---------------------------------------------------------------------------------------------------------
template <int n, bool EXT>
__global__ void __launch_bounds__(128, EXT ? 1 : 4) k_riemann(DevMesh m, Phys ph, int fBegin, int fEnd) {
    constexpr int N2 = n * n;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = fBegin + (int)(t / N2), mm = (int)(t % N2);
    if (f >= fEnd) return;
    const size_t fs = (size_t)m.nFace * N2, fo = (size_t)f * N2 + mm;
    const int info = m.faceInfo[f];
    const int ftype = info & 3, zone = (info >> 8) - 1;
    double nh[3], t1[3], t2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { nh[d] = m.fN[d * fs + fo]; t1[d] = 0.0; t2[d] = 0.0; }
    if (ph.riemann != H3D_RIEMANN_ROE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { t1[d] = m.fT1[d * fs + fo]; t2[d] = m.fT2[d * fs + fo]; }
    }
    const double Jf = m.fJ[fo];
    double QL[5], QR[5], visc[5] = {0, 0, 0, 0, 0}, inv[5];
    const double* fQ = m.fQ + fo; const double* fU = m.fU + fo;
    if (ftype == H3D_FACE_BOUNDARY) {
        const int btype = m.bcType[zone]; const double* P = m.bcParams + 16 * zone;
#pragma unroll
        for (int q = 0; q < 5; ++q) { QL[q] = fQ[q * fs]; QR[q] = QL[q]; }
        bc_flow_state(ph, btype, P, nh, QR);
        if (ph.ns) {
            double gx[5], gy[5], gz[5], F[5][3], mu, kappa;
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = fU[(size_t)(0 * 10 + q) * fs]; gy[q] = fU[(size_t)(1 * 10 + q) * fs]; gz[q] = fU[(size_t)(2 * 10 + q) * fs]; }
            laminar_mu_kappa(ph, QL, mu, kappa);
            if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<EXT>(ph, m.fDelta[f], ph.wallModel ? m.fDWall[fo] : 0.0, QL, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux<EXT>(ph, QL, gx, gy, gz, mu, 0.0, kappa, F);
#pragma unroll
            for (int q = 0; q < 5; ++q) { visc[q] = F[q][0] * nh[0] + F[q][1] * nh[1] + F[q][2] * nh[2]; }
            bc_neumann(btype, P, QL, visc);
        }
        riemann_solver<EXT>(ph, QL, QR, nh, t1, t2, inv);
    } else {
#pragma unroll
        for (int q = 0; q < 5; ++q) { QL[q] = fQ[q * fs]; QR[q] = fQ[(size_t)(5 + q) * fs]; }
        if (ph.ns) {
            double gx[5], gy[5], gz[5], FL[5][3], FR[5][3], mu, kappa;
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = fU[(size_t)(0 * 10 + q) * fs]; gy[q] = fU[(size_t)(1 * 10 + q) * fs]; gz[q] = fU[(size_t)(2 * 10 + q) * fs]; }
            laminar_mu_kappa(ph, QL, mu, kappa);
            if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<EXT>(ph, m.fDelta[f], ph.wallModel ? m.fDWall[fo] : 0.0, QL, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux<EXT>(ph, QL, gx, gy, gz, mu, 0.0, kappa, FL);
#pragma unroll
            for (int q = 0; q < 5; ++q) { gx[q] = fU[(size_t)(0 * 10 + 5 + q) * fs]; gy[q] = fU[(size_t)(1 * 10 + 5 + q) * fs]; gz[q] = fU[(size_t)(2 * 10 + 5 + q) * fs]; }
            laminar_mu_kappa(ph, QR, mu, kappa);
            if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<EXT>(ph, m.fDelta[f], ph.wallModel ? m.fDWall[fo] : 0.0, QR, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
            viscous_flux<EXT>(ph, QR, gx, gy, gz, mu, 0.0, kappa, FR);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const double fx = 0.5 * (FL[q][0] + FR[q][0]), fy = 0.5 * (FL[q][1] + FR[q][1]), fz = 0.5 * (FL[q][2] + FR[q][2]);
                visc[q] = fx * nh[0] + fy * nh[1] + fz * nh[2];
            }
            if (ph.viscous == H3D_VISCOUS_IP) {   // IP_RiemannSolver (EllipticIP.f90:704-761)
                const double penalty = ph.penaltyNum / m.fH[f];
#pragma unroll
                for (int q = 0; q < 5; ++q) visc[q] = visc[q] - penalty * ph.mu * (QL[q] - QR[q]);
            }
        }
        riemann_solver<EXT>(ph, QL, QR, nh, t1, t2, inv);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) m.fStar[(size_t)q * fs + fo] = (inv[q] - visc[q]) * Jf;
}

// ---------------------------------------------------------------------------------------------------------
// Volume term + surface lift + 1/J + source + RK update (+ fused prolongation of the updated state).
//   TimeDerivative_VolumetricContribution (SpatialDiscretization.f90:1602-1682)
//   BaseClass_ComputeInnerFluxes (HyperbolicDiscretizationClass.f90:83-152), BR1_ComputeInnerFluxes (EllipticBR1.f90:740-814)
//   ScalarWeakIntegrals_StdVolumeGreen (DGIntegrals.f90:56-87); SplitDG: HyperbolicSplitForm.f90:64-116 + DGIntegrals.f90:92-129
//   TimeDerivative_FacesContribution -> ScalarWeakIntegrals_StdFace (SpatialDiscretization.f90:1686-1701; DGIntegrals.f90:214-273)
//   QDot /= jacobian (:489-491); QDot += S_NS (:632-638); TakeRK3Step element loop (ExplicitMethods.f90:766-771)
// Shared memory: StandardDG: total contravariant flux [3][5][NS].  SplitDG: state [5][NS] + metrics [9][NS]
// (+ viscous contravariant flux [3][5][NS] for Navier-Stokes).  Plus fStar at element-trace nodes [6][5][N2].
// ---------------------------------------------------------------------------------------------------------
template <int n, bool TMA>
struct VolSmem {
    using C = KCfg<n>;
    __host__ __device__ static constexpr int fluxFields(bool split, bool ns) { return split ? (ns ? 15 : 0) : 15; }
    __host__ __device__ static constexpr int stagedFields(bool ns) { return TMA ? (ns ? 29 : 14) + 6 : 0; }   // Q 5 [, Ux Uy Uz 15], Ja 9 | J 1, G 5
    __host__ __device__ static constexpr int fields(bool split, bool ns) {
        return C::EPB * (stagedFields(ns) * C::N3 + (fluxFields(split, ns) + (split ? 15 : 0)) * C::NS + 30 * C::N2);   // split: Q 5, Ja 9, X 1
    }
    // face tables: one copy, or two (prefetched by bulk copies) in the staged variant; 4 mbarriers
    static size_t bytes(bool split, bool ns) { return sizeof(double) * (fields(split, ns) + 2 * C::N2 + 2 * n) + sizeof(int) * (TMA ? 2 : 1) * C::EPB * (6 * C::N2 + 8) + 48; }
};

// MODE 0: StandardDG; 1: SplitDG with the Kennedy-Gruber / Pirozzoli two-point fluxes; 2: SplitDG with all averages
// MMA (n = 8, StandardDG, staged): the three contractions run on the FP64 tensor cores (h3d_mma.cuh); the flux fields are then
// stored unpadded and swizzled, and the node-per-thread phase reads the contracted sums back from shared memory.
template <int n, int MODE, bool TMA, bool GV = false, bool MMA = false>
__global__ void __launch_bounds__(KCfg<n>::NT, KCfg<n>::MINB) k_volume(DevMesh m, Phys ph, RkArgs rk, const __grid_constant__ Ops<n> ops, int eBegin, int eEnd) {
    using C = KCfg<n>;
    constexpr bool SPLIT = MODE != 0, EXT = MODE == 2;
    static_assert(!MMA || (n == 8 && MODE == 0 && TMA && C::EPB == 1), "the DMMA contraction is written for the staged n = 8 StandardDG kernel");
    constexpr int N2 = C::N2, N3 = C::N3, NP = C::NP, NS = C::NS, EPB = C::EPB, TPE = C::TPE, NT = C::NT;
    constexpr int TN3 = EPB * N3;
    // Work fields (fluxes; SplitDG: state, metrics, sixth primitive) are stored UNPADDED, node = (k n + j) n + i: the lines the
    // contraction reads are then broadcasts (xi, eta) or consecutive words (zeta), free of bank conflicts; rows padded to an
    // odd length (round 1) cost the zeta lines a 3-way conflict and the whole phase twice the wavefronts
    // (profiles/r2_d_gen2/regions_gen2_core.txt: 94 M wavefronts against 200 M).  Only the prolongation buffer sP is padded.
    constexpr int WS = N3;
    constexpr int FSI = (EPB * 30 * N2 + NT - 1) / NT;   // fStar items per thread
    // MODE 1 stages HALVED primitives and metrics (two_point_flux_half).  Measured on B200, Euler Pirozzoli, volume kernel:
    // n=4 3.17 -> 3.11 ms, n=8 3.54 -> 3.35 ms, n=10 5.90 -> 5.66 ms, but n=6 3.23 -> 3.47 ms (three runs each): not at n=6.
    constexpr bool HALF = (MODE == 1) && (n != 6);
    // pair loop of the split form unrolled by LU: several pairs in flight give the FP64 pipe independent work.  Measured (Euler
    // Pirozzoli volume kernel, LU = 1 / 2 / 4): n=4 2.99 / 2.86 / 3.01 ms, n=6 3.03 / 3.02 / 3.35, n=8 3.37 / 3.20 / 3.14, n=10 5.66 /
    // 5.52 / 5.56; the general instantiation (MODE 2, logarithmic means) is slower unrolled (11.7 -> 12.5 ms) and stays rolled
    constexpr int LU = (MODE != 1) ? 1 : (n == 8 ? 4 : 2);
    extern __shared__ __align__(16) double smem[];
    const bool ns = ph.ns != 0;
    const int nStaged = TMA ? (ns ? 29 : 14) : 0;
    const int nFlux = SPLIT ? (ns ? 15 : 0) : 15;
    double* sIn = smem;                                  // TMA: [nStaged][TN3]: Q 5, (Ux Uy Uz 15,) Ja 9
    double* sJG = smem + nStaged * TN3;                  // TMA: [6][TN3]: J, G 5 (second barrier: consumed after the contraction)
    double* sF = sJG + (TMA ? 6 * TN3 : 0);              // [EPB][nFlux][WS]
    double* sQ = sF + EPB * nFlux * WS;                  // SPLIT: [EPB][5][WS]
    double* sJa = sQ + (SPLIT ? EPB * 5 * WS : 0);       // SPLIT: [EPB][9][WS]
    double* sX = sJa + (SPLIT ? EPB * 9 * WS : 0);       // SPLIT: [EPB][WS] sixth per-node primitive (see two_point_flux_prim)
    double* sFs = sF + EPB * (nFlux + (SPLIT ? 15 : 0)) * NS;   // [EPB][6][5][N2] fStar at element-trace nodes (signed); behind the padded extent
    double* sHatDT = sFs + EPB * 6 * 5 * N2;             // [n][n]
    double* sSharpDT = sHatDT + N2;                      // [n][n]
    double* sB = sSharpDT + N2;                          // [2][n]
    constexpr int TABI = EPB * (6 * N2 + 8);             // ints of one face-table set: trace offsets [EPB][6][N2] + info [EPB][8]
    int* sTab = (int*)(sB + 2 * n);                      // TMA: two sets, the next tile's is prefetched by bulk copies
    uint64_t* bar = (uint64_t*)(((uintptr_t)(sTab + (TMA ? 2 : 1) * TABI) + 7) & ~(uintptr_t)7);   // bar[0]: flux inputs, bar[1]: J and G, bar[2..3]: face tables
    double* sP = SPLIT ? sQ : sF;                        // prolongation buffer for the updated state [EPB][5][NS] (padded rows; SplitDG: runs into sJa)
    const int le = threadIdx.x / TPE, tn = threadIdx.x % TPE;
    const size_t es = (size_t)m.nElem * N3, fs = (size_t)m.nFace * N2;
    const int nTiles = (eEnd - eBegin + EPB - 1) / EPB;
    const int jaOff = ns ? 20 : 5;                       // first metric field inside sIn
    for (int t = threadIdx.x; t < N2; t += blockDim.x) { sHatDT[t] = m.hatDT[t]; if (SPLIT) sSharpDT[t] = m.sharpDT[t]; }
    if (threadIdx.x < 2 * n) sB[threadIdx.x] = m.b[threadIdx.x];
    auto issueTab = [&](int tile, int buf) {   // thread 0: face tables of the tile
        const int e0 = eBegin + tile * EPB;
        const int nLoc = min(EPB, eEnd - e0);
        int* dst = sTab + buf * TABI;
        mbar_arrive_expect_tx(bar + 2 + buf, (uint32_t)(nLoc * (6 * N2 + 8) * sizeof(int)));
        bulk_g2s(dst, m.elemTrace + (size_t)e0 * 6 * N2, (uint32_t)(nLoc * 6 * N2 * sizeof(int)), bar + 2 + buf);
        bulk_g2s(dst + EPB * 6 * N2, m.elemInfo + (size_t)e0 * 8, (uint32_t)(nLoc * 8 * sizeof(int)), bar + 2 + buf);
    };
    // (spreading these copies over the warps as k_gradient does costs this kernel 4 %: no warp waits for warp 0 here)
    // Warp 0 issues the bulk copies of a tile, lane c the copy of staged field c: one warp instruction instead of a loop of 29
    // in thread 0 (which kept warp 0 ~1.8 k cycles behind the others at the next barrier, profiles/r2_d_gen2/regions_gen1_mma.txt)
    auto issue = [&](int tile) {   // lanes of warp 0: bulk copies of the tile's staged fields
        const int e0 = eBegin + tile * EPB;
        const int nLoc = min(EPB, eEnd - e0);
        const uint32_t bytes = (uint32_t)(nLoc * N3 * sizeof(double));
        const size_t off = (size_t)e0 * N3;
        const int c = threadIdx.x;
        if (c == 0) mbar_arrive_expect_tx(bar, (uint32_t)nStaged * bytes);
        __syncwarp();
        if (c < nStaged) {
            const double* src;
            if (c < 5) src = m.Q + c * es;
            else if (c >= jaOff) src = m.Ja + (c - jaOff) * es;
            else { const int g = c - 5; src = (g < 5 ? m.Ux : (g < 10 ? m.Uy : m.Uz)) + (g % 5) * es; }
            bulk_g2s(sIn + c * TN3, src + off, bytes, bar);
        }
    };
    auto issueLate = [&](int tile) {   // lanes of warp 0: J and G of the tile (read after the contraction)
        const int e0 = eBegin + tile * EPB;
        const int nLoc = min(EPB, eEnd - e0);
        const uint32_t bytes = (uint32_t)(nLoc * N3 * sizeof(double));
        const size_t off = (size_t)e0 * N3;
        const int c = threadIdx.x;
        if (c == 0) mbar_arrive_expect_tx(bar + 1, 6 * bytes);
        __syncwarp();
        if (c < 6) bulk_g2s(sJG + c * TN3, (c == 0 ? m.J : m.G + (c - 1) * es) + off, bytes, bar + 1);
    };
    if (TMA) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); mbar_init(bar + 3, 1); fence_barrier_init(); }
        __syncthreads();
        if (threadIdx.x < 32 && (int)blockIdx.x < nTiles) { if (threadIdx.x == 0) issueTab(blockIdx.x, 0); issue(blockIdx.x); issueLate(blockIdx.x); }
    }
    uint32_t parity = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++iter) {
        const int e0 = eBegin + tile * EPB, e = e0 + le;
        const int nLocal = min(EPB, eEnd - e0);
        const bool active = e < eEnd;
        const int tbuf = TMA ? (iter & 1) : 0;
        int* sTr = sTab + tbuf * TABI;                   // [EPB][6][N2]
        int* sInfo = sTr + EPB * 6 * N2;                 // [EPB][8]
        if (TMA) {
            // this tile's tables were copied during the previous tile; the other set was last read by the previous
            // tile's prolongation, which the barrier at the end of the tile has closed
            if (threadIdx.x == 0 && tile + (int)gridDim.x < nTiles) issueTab(tile + gridDim.x, tbuf ^ 1);
            mbar_wait(bar + 2 + tbuf, (uint32_t)((iter >> 1) & 1));
        } else {
            load_face_tables<n>(m, sTr, sInfo, e0, nLocal);
            __syncthreads();
        }
        // interface fluxes of the six faces at element-trace nodes: raw loads issued now, signed (left +, right -,
        // FaceClass.f90:681-690) and stored to shared memory after the flux phase
        double fsv[FSI]; int fsg[FSI];
#pragma unroll
        for (int it = 0; it < FSI; ++it) {
            const int o = threadIdx.x + it * NT;
            fsv[it] = 0.0; fsg[it] = 0;
            if (o < nLocal * 30 * N2) {
                const int ab = o % N2; int r = o / N2;
                const int q = r % 5; r /= 5;                 // r = l2*6 + lf
                fsg[it] = sInfo[(r / 6) * 8 + r % 6];
                fsv[it] = m.fStar[(size_t)q * fs + sTr[r * N2 + ab]];
            }
        }
        if (TMA) { mbar_wait(bar, parity); parity ^= 1; }
        // SplitDG: per-node primitives are staged instead of the conserved state where the two-point flux allows it
        // (MODE 1 serves Kennedy-Gruber and Pirozzoli only and always does; MODE 2 decides at run time)
        const bool prim = (MODE == 1) || (EXT && prim_two_point_ok(ph.averaging));
        double Qk[5];                        // state of the thread's node (kept for the update)
        double Pk[SPLIT ? 6 : 1];            // SplitDG: primitives of the thread's node
        double FinvD[SPLIT ? 15 : 1];        // SplitDG: consistent (diagonal) inviscid contravariant fluxes
        if (active) {
            {
                const int node = tn;
                {
                    const int p = MMA ? swzF(node) : node;
                    const size_t go = (size_t)e * N3 + node;
                    const int so = le * N3 + node;
                    double ja[9];
                    if (TMA) {
#pragma unroll
                        for (int q = 0; q < 5; ++q) Qk[q] = sIn[q * TN3 + so];
#pragma unroll
                        for (int c = 0; c < 9; ++c) ja[c] = sIn[(jaOff + c) * TN3 + so];
                    } else {
#pragma unroll
                        for (int q = 0; q < 5; ++q) Qk[q] = m.Q[q * es + go];
#pragma unroll
                        for (int c = 0; c < 9; ++c) ja[c] = m.Ja[c * es + go];
                    }
                    // The viscous contravariant flux comes first: its inputs (15 gradient values) are dead before the inviscid
                    // flux is formed, which keeps the live set of this phase at ~45 doubles instead of ~60 (the kernel is bound
                    // by its 128 registers: a build with 16 more bytes of spills measured 20 % slower).
                    double F[5][3], fv[5][3];
                    if (ns) {
                        double gx[5], gy[5], gz[5], mu, kappa;
                        if (TMA) {
#pragma unroll
                            for (int q = 0; q < 5; ++q) { gx[q] = sIn[(5 + q) * TN3 + so]; gy[q] = sIn[(10 + q) * TN3 + so]; gz[q] = sIn[(15 + q) * TN3 + so]; }
                        } else {
#pragma unroll
                            for (int q = 0; q < 5; ++q) { gx[q] = m.Ux[q * es + go]; gy[q] = m.Uy[q * es + go]; gz[q] = m.Uz[q * es + go]; }
                        }
                        laminar_mu_kappa(ph, Qk, mu, kappa);
                        if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<GV>(ph, m.lesDelta[e], ph.wallModel ? m.dWall[go] : 0.0, Qk, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                        viscous_flux<GV>(ph, Qk, gx, gy, gz, mu, 0.0, kappa, F);
#pragma unroll
                        for (int q = 0; q < 5; ++q)
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                fv[q][d] = F[q][0] * ja[3 * d + 0] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
                                if (SPLIT) sF[((le * 3 + d) * 5 + q) * WS + p] = fv[q][d];
                            }
                    }
                    euler_flux(ph, Qk, F);
#pragma unroll
                    for (int q = 0; q < 5; ++q)
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const double fc = F[q][0] * ja[3 * d + 0] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
                            if (SPLIT) FinvD[d * 5 + q] = fc;
                            else sF[((le * 3 + d) * 5 + q) * WS + p] = fc - (ns ? fv[q][d] : 0.0);
                        }
                    if (SPLIT) {
                        if (prim) {
                            node_primitives(ph, Qk, Pk);
                            if (HALF) {   // halved primitives and metrics (two_point_flux_half)
#pragma unroll
                                for (int q = 0; q < 6; ++q) Pk[q] = 0.5 * Pk[q];
                            }
                            sX[le * WS + p] = Pk[5];
                        }
#pragma unroll
                        for (int q = 0; q < 5; ++q) sQ[(le * 5 + q) * WS + p] = prim ? Pk[q] : Qk[q];
#pragma unroll
                        for (int c = 0; c < 9; ++c) sJa[(le * 9 + c) * WS + p] = HALF ? 0.5 * ja[c] : ja[c];
                    }
                }
            }
        }
#pragma unroll
        for (int it = 0; it < FSI; ++it) {
            const int o = threadIdx.x + it * NT;
            if (o < nLocal * 30 * N2) sFs[o] = (fsg[it] & 1) ? -fsv[it] : fsv[it];     // o = ((l2*6 + lf)*5 + q)*N2 + ab
        }
        __syncthreads();
        if (TMA && threadIdx.x < 32 && tile + (int)gridDim.x < nTiles) issue(tile + gridDim.x);   // the staged inputs are consumed
        if constexpr (MMA) mma_volume_contract<NT / 32>(sF, sHatDT);
        if (TMA) mbar_wait(bar + 1, parity ^ 1);   // J and G of this tile (parity was flipped after the first wait)
        if (active) {
            {
                const int node = tn;
                {
                    const int i = node % n, j = (node / n) % n, k = node / N2;
                    const size_t go = (size_t)e * N3 + node;
                    const int bx = (k * n + j) * n, by = (k * n) * n + i, bz = j * n + i;
                    double vol[5] = {0, 0, 0, 0, 0};
                    if (MMA) {
#pragma unroll
                        for (int q = 0; q < 5; ++q) vol[q] = sF[q * WS + swzR(node)];
                    } else if (!SPLIT) {
                        const double* F1 = sF + ((le * 3 + 0) * 5) * WS + bx; const double* F2 = sF + ((le * 3 + 1) * 5) * WS + by; const double* F3 = sF + ((le * 3 + 2) * 5) * WS + bz;
#pragma unroll
                        for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + i];
#pragma unroll
                            for (int q = 0; q < 5; ++q) vol[q] = vol[q] + d * F1[q * WS + l]; }
#pragma unroll
                        for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + j];
#pragma unroll
                            for (int q = 0; q < 5; ++q) vol[q] = vol[q] + d * F2[q * WS + l * n]; }
#pragma unroll
                        for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + k];
#pragma unroll
                            for (int q = 0; q < 5; ++q) vol[q] = vol[q] + d * F3[q * WS + l * N2]; }
                    } else {
                        const double* sQe = sQ + le * 5 * WS; const double* sJe = sJa + le * 9 * WS; const double* sFe = sF + le * 15 * WS;
                        const int p = node;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const int me = d == 0 ? i : (d == 1 ? j : k);
                            const int ob = d == 0 ? bx : (d == 1 ? by : bz);
                            const int ostr = d == 0 ? 1 : (d == 1 ? n : N2);
                            double jaMe[3];
#pragma unroll
                            for (int c = 0; c < 3; ++c) jaMe[c] = sJe[(3 * d + c) * WS + p];
#pragma unroll LU
                            for (int l = 0; l < n; ++l) {
                                const int other = ob + l * ostr;
                                double fsvv[5];
                                if (l == me) {
#pragma unroll
                                    for (int q = 0; q < 5; ++q) fsvv[q] = FinvD[d * 5 + q];
                                } else {
                                    double Qo[6], jo[3];
#pragma unroll
                                    for (int q = 0; q < 5; ++q) Qo[q] = sQe[q * WS + other];
#pragma unroll
                                    for (int c = 0; c < 3; ++c) jo[c] = sJe[(3 * d + c) * WS + other];
                                    if (prim) {
                                        Qo[5] = sX[le * WS + other];
                                        if (HALF) { if (l > me) two_point_flux_half(ph, Pk, Qo, jaMe, jo, fsvv); else two_point_flux_half(ph, Qo, Pk, jo, jaMe, fsvv); }
                                        else if (l > me) two_point_flux_prim<EXT>(ph, Pk, Qo, jaMe, jo, fsvv); else two_point_flux_prim<EXT>(ph, Qo, Pk, jo, jaMe, fsvv);
                                    } else {
                                        if (l > me) two_point_flux<EXT>(ph, Qk, Qo, jaMe, jo, fsvv); else two_point_flux<EXT>(ph, Qo, Qk, jo, jaMe, fsvv);
                                    }
                                }
                                const double sd = sSharpDT[l * n + me], hd = sHatDT[l * n + me];
                                if (ns) {
#pragma unroll
                                    for (int q = 0; q < 5; ++q) vol[q] = vol[q] + sd * fsvv[q] + hd * sFe[(d * 5 + q) * WS + other];
                                } else {
#pragma unroll
                                    for (int q = 0; q < 5; ++q) vol[q] = vol[q] + sd * fsvv[q];
                                }
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 5; ++q) vol[q] = -vol[q];
                    }
                    // surface integral in the reference's order L,R,FRONT,BACK,BOTTOM,TOP
                    const double* Fs = sFs + (le * 6) * 5 * N2;
                    const double bL = sB[i], bR = sB[n + i], bF = sB[j], bBk = sB[n + j], bBo = sB[k], bT = sB[n + k];
                    double Jn, Gk[5];
                    if (TMA) {
                        Jn = sJG[le * N3 + node];
#pragma unroll
                        for (int q = 0; q < 5; ++q) Gk[q] = sJG[(1 + q) * TN3 + le * N3 + node];
                    } else {
                        Jn = m.J[go];
#pragma unroll
                        for (int q = 0; q < 5; ++q) Gk[q] = (rk.mode != 0) ? m.G[q * es + go] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        double fi = Fs[(5 * 5 + q) * N2 + k * n + j] * bL;
                        fi = fi + Fs[(3 * 5 + q) * N2 + k * n + j] * bR;
                        fi = fi + Fs[(0 * 5 + q) * N2 + k * n + i] * bF;
                        fi = fi + Fs[(1 * 5 + q) * N2 + k * n + i] * bBk;
                        fi = fi + Fs[(2 * 5 + q) * N2 + j * n + i] * bBo;
                        fi = fi + Fs[(4 * 5 + q) * N2 + j * n + i] * bT;
                        double res = vol[q] - fi;
                        res = res / Jn;
                        if (m.S) res = res + m.S[q * es + go];
                        if (rk.mode == 0) {
                            m.QDot[q * es + go] = res;
                        } else if (rk.mode == 1) {
                            if (rk.storeQDot) m.QDot[q * es + go] = res;
                            const double gg = rk.a * Gk[q] + res;
                            m.G[q * es + go] = gg;
                            Qk[q] = Qk[q] + rk.cdt * gg;
                            m.Q[q * es + go] = Qk[q];
                        } else {   // TakeSSPRK33Step / TakeSSPRK43Step (ExplicitMethods.f90:983-1230)
                            if (rk.storeQDot) m.QDot[q * es + go] = res;
                            const double g0 = rk.copyG ? Qk[q] : Gk[q];
                            if (rk.copyG) m.G[q * es + go] = g0;
                            Qk[q] = rk.a * g0 + rk.b * Qk[q] + rk.cdt * res;
                            m.Q[q * es + go] = Qk[q];
                        }
                    }
                }
            }
        }
        if (rk.prolong) {
            __syncthreads();
            if (active) {
                {
                    const int node = tn;
                    {
#pragma unroll
                        for (int q = 0; q < 5; ++q) sP[(le * 5 + q) * NS + C::pidx(node)] = Qk[q];
                    }
                }
            }
            __syncthreads();
            prolong_block<n, 5>(m, ops, sP, sTr, sInfo, m.fQ, nLocal);
        }
        __syncthreads();
        if (TMA && threadIdx.x < 32 && tile + (int)gridDim.x < nTiles) issueLate(tile + gridDim.x);   // J/G buffer is free again
    }
}

// shared-memory footprints (bytes)
template <int n> inline size_t smemProlong() { using C = KCfg<n>; return sizeof(double) * ((size_t)C::EPB * 5 * C::NS + 2 * n) + sizeof(int) * C::EPB * (6 * C::N2 + 8); }
template <int n, bool TMA, bool VISC = false> inline size_t smemGradient() { return GradSmem<n, TMA, VISC>::bytes; }
template <int n, bool TMA> inline size_t smemVolume(bool split, bool ns) { return VolSmem<n, TMA>::bytes(split, ns); }

}  // namespace h3d
