// FP64 tensor-core (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) contractions of one n = 8 element with the derivative matrices.
//
// The north star asks for the experiment "FP64 tensor-core MMA only if ncu shows it beats the CUDA-core path".  Measured on the
// B200 (scripts/micro/dmma_rate.cu, profiles/r2_a_parity_dmma/dmma_rate.txt): 0.25 DMMA / cycle / SM = 64 MAC / cycle / SM at a
// 26-cycle latency, against 32 MAC / cycle / SM of the DMUL+DADD pairs of the bit-exact build -- and, what matters more, the
// operands of a DMMA are two registers per lane, where the node-per-thread contraction reads n values per direction, equation
// and node from shared memory (120 LDS.64 per node; the phase is bound by shared-memory wavefronts).
//
// One DMMA is C[8][8] += A[8][4] B[4][8].  Lane L = 4 g + a (g = L / 4, a = L % 4) holds A[g][a], B[a][g], C[g][2a], C[g][2a+1].
// A contraction along an element axis is (8 x 8 operator) x (8 x 64 lines) per equation: 8 column blocks x 2 k-halves = 16 DMMAs
// per equation and direction, 240 per element for the 15 (direction, equation) pairs.
//
//   xi   (tile = plane k, rows j, columns i'): C[j][i'] = sum_i F(i,j,k) M(i',i)       A = field  (one LDS.128: i = 2a, 2a+1), B = operator
//   eta  (tile = plane k, rows j', columns i): C[j'][i] = sum_j M(j',j) F(i,j,k)       A = operator, B = field (j = a, a + 4)
//   zeta (tile = plane j, rows k', columns i): C[k'][i] = sum_k M(k',k) F(i,j,k)       A = operator, B = field (k = a, a + 4)
//
// xi and eta of the volume term accumulate into the same C registers (both tiles are the plane k in the same layout); zeta lives
// on another set of planes and is added through shared memory.  The summation is fused and in another order than the reference's,
// so this path is NOT bit-identical to the oracle (1e-13..1e-12 of the field's max-norm, asserted by the tests); the CUDA-core
// contraction stays selectable (h3d_set_option "mma=0").
#pragma once
#include <cuda_runtime.h>

namespace h3d {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Shared-memory layouts of one 512-node field, node = i + 8 j + 64 k, unpadded and XOR-swizzled so that every fragment access
// below is free of bank conflicts (bits 2-3 of the index are flipped by bits of j and k; 16-double blocks stay in place):
//   flux fields : node ^ 4 * (((j >> 1) & 1) ^ (k & 3))
//   result field: node ^ 8 * (k & 1)
__device__ __forceinline__ int swzF(int node) { return node ^ ((((node >> 4) & 1) ^ ((node >> 6) & 3)) << 2); }
__device__ __forceinline__ int swzR(int node) { return node ^ (((node >> 6) & 1) << 3); }

// Volume term: R[q][node] = sum_l hatD(i,l) F1_q(l,j,k) + sum_l hatD(j,l) F2_q(i,l,k) + sum_l hatD(k,l) F3_q(i,j,l)
// (ScalarWeakIntegrals_StdVolumeGreen, DGIntegrals.f90:56-87).  sF: [15][512] = (direction, equation) fields in the swzF layout;
// the result of equation q overwrites F1_q in the swzR layout.  sMT[l*8 + i] = hatD(i,l).  All threads of the CTA call it.
// Every warp keeps its UPW = ceil(40 / NWARPS) tiles in flight at once (independent accumulators; a tile alone is a chain of
// LDS -> 4 dependent DMMAs of 26 cycles each -> STS); a slot beyond the 40 tiles repeats tile 39 and drops its result.
template <int NWARPS>
__device__ __forceinline__ void mma_volume_contract(double* __restrict__ sF, const double* __restrict__ sMT) {
    constexpr int UPW = (40 + NWARPS - 1) / NWARPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, a = lane & 3;
    const double bx0 = sMT[(2 * a) * 8 + g], bx1 = sMT[(2 * a + 1) * 8 + g];   // B of the xi tiles: rows l = 2a + h, column i' = g
    const double ay0 = sMT[a * 8 + g], ay1 = sMT[(a + 4) * 8 + g];             // A of the eta / zeta tiles: row g, columns l = a + 4 h
    {
        double2 f1[UPW]; double f20[UPW], f21[UPW], c0[UPW], c1[UPW];
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const int u = min(warp + s * NWARPS, 39), k = u / 5, q = u - 5 * k;
            f1[s] = *reinterpret_cast<const double2*>(sF + q * 512 + swzF(k * 64 + g * 8 + 2 * a));
            const double* F2 = sF + (5 + q) * 512;
            f20[s] = F2[swzF(k * 64 + a * 8 + g)]; f21[s] = F2[swzF(k * 64 + (a + 4) * 8 + g)];
            c0[s] = 0.0; c1[s] = 0.0;
        }
#pragma unroll
        for (int s = 0; s < UPW; ++s) dmma884(c0[s], c1[s], f1[s].x, bx0);
#pragma unroll
        for (int s = 0; s < UPW; ++s) dmma884(c0[s], c1[s], f1[s].y, bx1);
#pragma unroll
        for (int s = 0; s < UPW; ++s) dmma884(c0[s], c1[s], ay0, f20[s]);
#pragma unroll
        for (int s = 0; s < UPW; ++s) dmma884(c0[s], c1[s], ay1, f21[s]);
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const int u = warp + s * NWARPS, k = u / 5, q = u - 5 * k;
            if (u < 40) *reinterpret_cast<double2*>(sF + q * 512 + swzR(k * 64 + g * 8 + 2 * a)) = make_double2(c0[s], c1[s]);
        }
    }
    __syncthreads();
    {
        double2 c[UPW]; double f30[UPW], f31[UPW];
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const int u = min(warp + s * NWARPS, 39), j = u / 5, q = u - 5 * j;
            c[s] = *reinterpret_cast<const double2*>(sF + q * 512 + swzR(g * 64 + j * 8 + 2 * a));
            const double* F3 = sF + (10 + q) * 512;
            f30[s] = F3[swzF(a * 64 + j * 8 + g)]; f31[s] = F3[swzF((a + 4) * 64 + j * 8 + g)];
        }
#pragma unroll
        for (int s = 0; s < UPW; ++s) dmma884(c[s].x, c[s].y, ay0, f30[s]);
#pragma unroll
        for (int s = 0; s < UPW; ++s) dmma884(c[s].x, c[s].y, ay1, f31[s]);
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const int u = warp + s * NWARPS, j = u / 5, q = u - 5 * j;
            if (u < 40) *reinterpret_cast<double2*>(sF + q * 512 + swzR(g * 64 + j * 8 + 2 * a)) = c[s];
        }
    }
    __syncthreads();
}

// Local gradient: G[(d*5 + q)][node] = sum_l D(node_d, l) U_q(.. l ..) for d = xi, eta, zeta
// (HexElement_ComputeLocalGradient, HexElementClass.f90:484-500).  sU: [5] fields of stride US, unpadded and unswizzled (staged by
// bulk copies); sG: [15] fields of stride GS, rows padded to 9 doubles (PSWZ false) or in the pswz layout of h3d_kernels2.cuh.
// sDT[l*8 + i] = D(i,l).  Per direction every warp keeps its ceil(40 / NWARPS) tiles in flight (see mma_volume_contract).
template <int NWARPS, int US, int GS, bool PSWZ = false>
__device__ __forceinline__ void mma_gradient_contract(const double* __restrict__ sU, double* __restrict__ sG, const double* __restrict__ sDT) {
    constexpr int UPW = (40 + NWARPS - 1) / NWARPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, a = lane & 3;
    const double bx0 = sDT[(2 * a) * 8 + g], bx1 = sDT[(2 * a + 1) * 8 + g];
    const double ay0 = sDT[a * 8 + g], ay1 = sDT[(a + 4) * 8 + g];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double f0[UPW], f1[UPW], c0[UPW], c1[UPW];
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const int u = min(warp + s * NWARPS, 39), pl = u / 5, q = u - 5 * pl;   // plane (k for xi / eta, j for zeta), equation
            const double* U = sU + q * US;
            if (d == 0) { const double2 f = *reinterpret_cast<const double2*>(U + pl * 64 + g * 8 + 2 * a); f0[s] = f.x; f1[s] = f.y; }
            else if (d == 1) { f0[s] = U[pl * 64 + a * 8 + g]; f1[s] = U[pl * 64 + (a + 4) * 8 + g]; }
            else { f0[s] = U[a * 64 + pl * 8 + g]; f1[s] = U[(a + 4) * 64 + pl * 8 + g]; }
            c0[s] = 0.0; c1[s] = 0.0;
        }
#pragma unroll
        for (int s = 0; s < UPW; ++s) { if (d == 0) dmma884(c0[s], c1[s], f0[s], bx0); else dmma884(c0[s], c1[s], ay0, f0[s]); }
#pragma unroll
        for (int s = 0; s < UPW; ++s) { if (d == 0) dmma884(c0[s], c1[s], f1[s], bx1); else dmma884(c0[s], c1[s], ay1, f1[s]); }
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const int u = warp + s * NWARPS, pl = u / 5, q = u - 5 * pl;
            const int node = d == 2 ? g * 64 + pl * 8 + 2 * a : pl * 64 + g * 8 + 2 * a;
            if (u < 40) {
                if (PSWZ) {   // node and node + 1 differ in bit 0 of i only: the swizzled positions are the two halves of one 16-byte pair
                    const int p = (node & ~63) | (((((node >> 3) & 7) ^ ((node >> 6) & 1))) << 3) | ((node & 7) ^ ((node >> 3) & 7));
                    double* o = sG + (d * 5 + q) * GS;
                    o[p] = c0[s]; o[p ^ 1] = c1[s];
                } else {
                    double* o = sG + (d * 5 + q) * GS + (node >> 3) * 9 + (node & 7);
                    o[0] = c0[s]; o[1] = c1[s];
                }
            }
        }
    }
    __syncthreads();
}

}  // namespace h3d
