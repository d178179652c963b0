// Second-generation element kernels for n = 8 (P = 7, the headline order): 256-thread CTAs, two nodes per thread, at most
// 112 KB of shared memory, so that TWO CTAs share an SM.
//
// Why (profiles/r1_h_final/regions.txt, profiles/r2_b_dmma_first): the first-generation kernels run one 512-thread CTA per
// SM whose phases add up -- the flux phase is bound by the FP64 pipe (shared memory idle), the contraction by shared-memory
// wavefronts (FP64 pipe at a third), and the bulk copies of the next tile can only start once the single staging buffer has been
// consumed, so no load is in flight during the flux phase.  With two independent CTAs per SM one CTA's flux phase runs beside
// the other's contraction or beside its wait for the next tile; each pipe then sees the sum of both CTAs' demands instead of the
// sum of all phases.
//
// What changes against k_volume / k_gradient (h3d_kernels.cuh):
//   * only the fields that need a prefetch distance are staged by bulk copies: grad U (15) and Q (5) in the volume kernel, Q (5)
//     in the gradient kernel; metrics, J and G are read with plain coalesced loads issued a phase ahead of their use;
//   * the contravariant fluxes overwrite the staged gradients in place (a node's 15 flux values take the slots of its 15
//     gradient values), the updated state overwrites the staged state: no separate work buffers;
//   * fields that are read along lines (prolongation) are XOR-swizzled instead of padded (pswz): same conflict-free access,
//     no 12 % padding, 16-byte alignment kept.
// Arithmetic, summation order and results are those of the first-generation kernels: bit-identical to the oracle (the DMMA
// variants excepted, see h3d_mma.cuh).
#pragma once
#include "h3d_kernels.cuh"

namespace h3d {

// swizzled position of node (i,j,k) = i + 8 j + 64 k inside a 512-node field: lines along any axis, read by the 64 threads of a
// face or by 16 consecutive nodes, fall into 16 distinct 8-byte banks
__device__ __forceinline__ int pswz(int node) {
    const int i = node & 7, j = (node >> 3) & 7, k = node >> 6;
    return (k << 6) | ((j ^ (k & 1)) << 3) | (i ^ j);
}

// prolongation of NV swizzled fields sF[v][pswz(node)] of ONE element to its six faces (cf. prolong_axis): 256 threads, thread
// (row, ab) takes the trace node ab of the fields row, row + 4, ...
template <int NV, int AX>
__device__ __forceinline__ void prolong2_axis(const DevMesh& m, const Ops<8>& ops, const double* __restrict__ sF, const int* __restrict__ sTr,
                                              const int* __restrict__ sInfo, double* __restrict__ dst) {
    constexpr int n = 8, N2 = 64;
    constexpr int LF0 = AX == 0 ? 5 : (AX == 1 ? 0 : 2), LF1 = AX == 0 ? 3 : (AX == 1 ? 1 : 4);   // LEFT,RIGHT | FRONT,BACK | BOTTOM,TOP
    const size_t fstride = (size_t)m.nFace * N2;
    const int ab = threadIdx.x & 63, a = ab & 7, b = ab >> 3;
    double* p0 = dst + (size_t)((sInfo[LF0] & 1) * 5) * fstride + sTr[LF0 * N2 + ab];
    double* p1 = dst + (size_t)((sInfo[LF1] & 1) * 5) * fstride + sTr[LF1 * N2 + ab];
#pragma unroll 1
    for (int vv = threadIdx.x >> 6; vv < NV; vv += 4) {
        const double* src = sF + vv * 512;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int l = 0; l < n; ++l) {
            const int node = AX == 0 ? (b * n + a) * n + l : (AX == 1 ? (b * n + l) * n + a : (l * n + b) * n + a);
            const double sv = src[pswz(node)];
            acc0 = acc0 + sv * ops.v[l]; acc1 = acc1 + sv * ops.v[n + l];
        }
        const size_t fo = (size_t)((vv / 5) * 10 + vv % 5) * fstride;
        p0[fo] = acc0; p1[fo] = acc1;
    }
}
template <int NV>
__device__ __forceinline__ void prolong2_block(const DevMesh& m, const Ops<8>& ops, const double* __restrict__ sF, const int* __restrict__ sTr,
                                               const int* __restrict__ sInfo, double* __restrict__ dst) {
    prolong2_axis<NV, 0>(m, ops, sF, sTr, sInfo, dst);
    prolong2_axis<NV, 1>(m, ops, sF, sTr, sInfo, dst);
    prolong2_axis<NV, 2>(m, ops, sF, sTr, sInfo, dst);
}

struct Vol2Smem {
    // grad U / fluxes [15][512], Q [5][512], fStar [6][5][64], hatD^T [64], b [16] | two face-table sets | 4 mbarriers
    static constexpr int doubles = 15 * 512 + 5 * 512 + 30 * 64 + 64 + 16;
    static constexpr int tabInts = 6 * 64 + 8;
    static constexpr size_t bytes = sizeof(double) * doubles + sizeof(int) * 2 * tabInts + 64;
};

// ---------------------------------------------------------------------------------------------------------
// Volume term + surface lift + 1/J + source + RK update + prolongation of the updated state, StandardDG, n = 8.
// Same references as k_volume.  MMA: contractions on the FP64 tensor cores.
// ---------------------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(256, 2) k_volume2(DevMesh m, Phys ph, RkArgs rk, const __grid_constant__ Ops<8> ops, int eBegin, int eEnd) {
    constexpr int n = 8, N2 = 64, N3 = 512, NT = 256;
    constexpr int FSI = (30 * N2 + NT - 1) / NT;   // fStar items per thread
    extern __shared__ __align__(16) double smem[];
    const bool ns = ph.ns != 0;
    double* sGU = smem;                         // [15][512]: grad U (bulk copies), then the contravariant fluxes in place
    double* sQ = sGU + 15 * N3;                 // [5][512]: Q (bulk copies), then the updated state (swizzled) for the prolongation
    double* sFs = sQ + 5 * N3;                  // [6][5][64] fStar at element-trace nodes (signed)
    double* sHatDT = sFs + 30 * N2;             // [8][8]
    double* sB = sHatDT + N2;                   // [2][8]
    constexpr int TABI = Vol2Smem::tabInts;
    int* sTab = (int*)(sB + 2 * n);
    uint64_t* bar = (uint64_t*)(((uintptr_t)(sTab + 2 * TABI) + 7) & ~(uintptr_t)7);   // bar[0]: grad U, bar[1]: Q, bar[2..3]: face tables
    const size_t es = (size_t)m.nElem * N3, fs = (size_t)m.nFace * N2;
    const int nTiles = eEnd - eBegin;
    const int tid = threadIdx.x;
    if (tid < N2) sHatDT[tid] = m.hatDT[tid];
    if (tid < 2 * n) sB[tid] = m.b[tid];
    auto issueTab = [&](int tile, int buf) {
        const int e0 = eBegin + tile;
        int* dst = sTab + buf * TABI;
        mbar_arrive_expect_tx(bar + 2 + buf, (uint32_t)(TABI * sizeof(int)));
        bulk_g2s(dst, m.elemTrace + (size_t)e0 * 6 * N2, (uint32_t)(6 * N2 * sizeof(int)), bar + 2 + buf);
        bulk_g2s(dst + 6 * N2, m.elemInfo + (size_t)e0 * 8, (uint32_t)(8 * sizeof(int)), bar + 2 + buf);
    };
    auto issueGU = [&](int tile) {
        const size_t off = (size_t)(eBegin + tile) * N3;
        fence_proxy_async();   // the buffer was written by ordinary stores (fluxes in place) before the barrier that released it
        mbar_arrive_expect_tx(bar, 15u * N3 * sizeof(double));
#pragma unroll 1
        for (int c = 0; c < 5; ++c) {
            bulk_g2s(sGU + c * N3, m.Ux + c * es + off, N3 * sizeof(double), bar);
            bulk_g2s(sGU + (5 + c) * N3, m.Uy + c * es + off, N3 * sizeof(double), bar);
            bulk_g2s(sGU + (10 + c) * N3, m.Uz + c * es + off, N3 * sizeof(double), bar);
        }
    };
    auto prefetchPlain = [&](int tile) {   // the fields read with plain loads (metrics, J, G): into L2 one tile ahead
        const size_t off = (size_t)(eBegin + tile) * N3;
#pragma unroll 1
        for (int c = 0; c < 9; ++c) bulk_prefetch_l2(m.Ja + c * es + off, N3 * sizeof(double));
        bulk_prefetch_l2(m.J + off, N3 * sizeof(double));
        if (rk.mode != 0) {
#pragma unroll 1
            for (int c = 0; c < 5; ++c) bulk_prefetch_l2(m.G + c * es + off, N3 * sizeof(double));
        }
    };
    auto issueQ = [&](int tile) {
        const size_t off = (size_t)(eBegin + tile) * N3;
        fence_proxy_async();
        mbar_arrive_expect_tx(bar + 1, 5u * N3 * sizeof(double));
#pragma unroll 1
        for (int c = 0; c < 5; ++c) bulk_g2s(sQ + c * N3, m.Q + c * es + off, N3 * sizeof(double), bar + 1);
    };
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); mbar_init(bar + 3, 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < nTiles) { issueTab(blockIdx.x, 0); if (ns) issueGU(blockIdx.x); issueQ(blockIdx.x); prefetchPlain(blockIdx.x); }
    uint32_t parity = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++iter) {
        const int e = eBegin + tile;
        const int next = tile + (int)gridDim.x;
        const int tbuf = iter & 1;
        const int* sTr = sTab + tbuf * TABI;
        const int* sInfo = sTr + 6 * N2;
        if (tid == 0 && next < nTiles) { issueTab(next, tbuf ^ 1); prefetchPlain(next); }   // the other table set was last read by the previous tile's prolongation
        mbar_wait(bar + 2 + tbuf, (uint32_t)((iter >> 1) & 1));
        // interface fluxes of the six faces at element-trace nodes, signed (left +, right -, FaceClass.f90:681-690)
        {
            double fsv[FSI];
#pragma unroll
            for (int it = 0; it < FSI; ++it) {
                const int o = tid + it * NT;
                fsv[it] = 0.0;
                if (o < 30 * N2) {
                    const int ab = o & 63, r = o >> 6, q = r % 5, lf = r / 5;
                    const double v = m.fStar[(size_t)q * fs + sTr[lf * N2 + ab]];
                    fsv[it] = (sInfo[lf] & 1) ? -v : v;
                }
            }
#pragma unroll
            for (int it = 0; it < FSI; ++it) {
                const int o = tid + it * NT;
                if (o < 30 * N2) sFs[o] = fsv[it];      // o = (lf*5 + q)*N2 + ab
            }
        }
        mbar_wait(bar + 1, parity);
        if (ns) mbar_wait(bar, parity);
        parity ^= 1;
        // ---- flux phase: contravariant fluxes of the thread's two nodes, written over the staged gradients
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            const int node = tid + r * NT;
            const size_t go = (size_t)e * N3 + node;
            double ja[9], Qk[5];
#pragma unroll
            for (int c = 0; c < 9; ++c) ja[c] = m.Ja[c * es + go];
#pragma unroll
            for (int q = 0; q < 5; ++q) Qk[q] = sQ[q * N3 + node];
            double F[5][3], fv[5][3];
            if (ns) {
                double gx[5], gy[5], gz[5], mu, kappa;
#pragma unroll
                for (int q = 0; q < 5; ++q) { gx[q] = sGU[q * N3 + node]; gy[q] = sGU[(5 + q) * N3 + node]; gz[q] = sGU[(10 + q) * N3 + node]; }
                laminar_mu_kappa(ph, Qk, mu, kappa);
                if (ph.les != H3D_LES_NONE) { const double mut = smagorinsky<false>(ph, m.lesDelta[e], ph.wallModel ? m.dWall[go] : 0.0, Qk, gx, gy, gz); mu = mu + mut; kappa = kappa + mut * ph.mu_to_kappa; }
                viscous_flux<false>(ph, Qk, gx, gy, gz, mu, 0.0, kappa, F);
#pragma unroll
                for (int q = 0; q < 5; ++q)
#pragma unroll
                    for (int d = 0; d < 3; ++d) fv[q][d] = F[q][0] * ja[3 * d + 0] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
            }
            euler_flux(ph, Qk, F);
            if (MMA) __syncwarp();   // the swizzle moves a node's fluxes into another lane's gradient slots (same half-warp)
            const int p = MMA ? swzF(node) : node;
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double fc = F[q][0] * ja[3 * d + 0] + F[q][1] * ja[3 * d + 1] + F[q][2] * ja[3 * d + 2];
                    sGU[(d * 5 + q) * N3 + p] = fc - (ns ? fv[q][d] : 0.0);
                }
        }
        __syncthreads();
        // ---- contraction (ScalarWeakIntegrals_StdVolumeGreen)
        double vol[2][5];
        if constexpr (MMA) {
            mma_volume_contract<NT / 32>(sGU, sHatDT);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < 5; ++q) vol[r][q] = sGU[q * N3 + swzR(tid + r * NT)];
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int node = tid + r * NT;
                const int i = node & 7, j = (node >> 3) & 7, k = node >> 6;
                const double* F1 = sGU + (k * n + j) * n; const double* F2 = sGU + 5 * N3 + (k * n) * n + i; const double* F3 = sGU + 10 * N3 + j * n + i;
#pragma unroll
                for (int q = 0; q < 5; ++q) vol[r][q] = 0.0;
#pragma unroll
                for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + i];
#pragma unroll
                    for (int q = 0; q < 5; ++q) vol[r][q] = vol[r][q] + d * F1[q * N3 + l]; }
#pragma unroll
                for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + j];
#pragma unroll
                    for (int q = 0; q < 5; ++q) vol[r][q] = vol[r][q] + d * F2[q * N3 + l * n]; }
#pragma unroll
                for (int l = 0; l < n; ++l) { const double d = sHatDT[l * n + k];
#pragma unroll
                    for (int q = 0; q < 5; ++q) vol[r][q] = vol[r][q] + d * F3[q * N3 + l * N2]; }
            }
        }
        __syncthreads();                                        // the fluxes are consumed
        if (tid == 0 && next < nTiles && ns) issueGU(next);
        // ---- surface integral (order L,R,FRONT,BACK,BOTTOM,TOP), 1/J, source, RK update; J and G come from L2 (prefetched)
        double Qn[2][5];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int node = tid + r * NT;
            const int i = node & 7, j = (node >> 3) & 7, k = node >> 6;
            const size_t go = (size_t)e * N3 + node;
            double Jn[2], Gk[2][5];
            Jn[r] = m.J[go];
#pragma unroll
            for (int q = 0; q < 5; ++q) Gk[r][q] = (rk.mode != 0) ? m.G[q * es + go] : 0.0;
            const double bL = sB[i], bR = sB[n + i], bF = sB[j], bBk = sB[n + j], bBo = sB[k], bT = sB[n + k];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                double fi = sFs[(5 * 5 + q) * N2 + k * n + j] * bL;
                fi = fi + sFs[(3 * 5 + q) * N2 + k * n + j] * bR;
                fi = fi + sFs[(0 * 5 + q) * N2 + k * n + i] * bF;
                fi = fi + sFs[(1 * 5 + q) * N2 + k * n + i] * bBk;
                fi = fi + sFs[(2 * 5 + q) * N2 + j * n + i] * bBo;
                fi = fi + sFs[(4 * 5 + q) * N2 + j * n + i] * bT;
                double res = vol[r][q] - fi;
                res = res / Jn[r];
                if (m.S) res = res + m.S[q * es + go];
                const double q0 = sQ[q * N3 + node];
                if (rk.mode == 0) {
                    m.QDot[q * es + go] = res;
                    Qn[r][q] = q0;
                } else if (rk.mode == 1) {
                    if (rk.storeQDot) m.QDot[q * es + go] = res;
                    const double gg = rk.a * Gk[r][q] + res;
                    m.G[q * es + go] = gg;
                    Qn[r][q] = q0 + rk.cdt * gg;
                    m.Q[q * es + go] = Qn[r][q];
                } else {   // TakeSSPRK33Step / TakeSSPRK43Step (ExplicitMethods.f90:983-1230)
                    if (rk.storeQDot) m.QDot[q * es + go] = res;
                    const double g0 = rk.copyG ? q0 : Gk[r][q];
                    if (rk.copyG) m.G[q * es + go] = g0;
                    Qn[r][q] = rk.a * g0 + rk.b * q0 + rk.cdt * res;
                    m.Q[q * es + go] = Qn[r][q];
                }
            }
        }
        __syncthreads();                                        // the staged state is consumed
        if (rk.prolong) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int p = pswz(tid + r * NT);
#pragma unroll
                for (int q = 0; q < 5; ++q) sQ[q * N3 + p] = Qn[r][q];
            }
            __syncthreads();
            prolong2_block<5>(m, ops, sQ, sTr, sInfo, m.fQ);
            __syncthreads();
        }
        if (tid == 0 && next < nTiles) issueQ(next);
    }
}

struct Grad2Smem {
    // Q [5][512], interface uStar [6][5][64] + normal, J_f [6][4][64], gradients [15][512], D^T [64], b [16] | two table sets | mbarriers
    static constexpr int doubles = 5 * 512 + 30 * 64 + 24 * 64 + 15 * 512 + 64 + 16;
    static constexpr int tabInts = 6 * 64 + 8;
    static constexpr size_t bytes = sizeof(double) * doubles + sizeof(int) * 2 * tabInts + 64;
};

// ---------------------------------------------------------------------------------------------------------
// BR1 gradient (local gradient + interface lift + prolongation of the gradients), gradient variables = state, n = 8.
// Same references as k_gradient.  MMA: the three derivative contractions on the FP64 tensor cores.
// ---------------------------------------------------------------------------------------------------------
template <bool MMA>
__global__ void __launch_bounds__(256, 2) k_gradient2(DevMesh m, Phys ph, const __grid_constant__ Ops<8> ops, int eBegin, int eEnd) {
    constexpr int n = 8, N2 = 64, N3 = 512, NT = 256;
    constexpr int IFI = (6 * N2 + NT - 1) / NT;   // interface items per thread
    extern __shared__ __align__(16) double smem[];
    double* sQ = smem;                          // [5][512] state (bulk copies)
    double* sH = sQ + 5 * N3;                   // [6][5][64]
    double* sNrm = sH + 30 * N2;                // [6][4][64]
    double* sG = sNrm + 24 * N2;                // [15][512] gradients in the pswz layout
    double* sDT = sG + 15 * N3;                 // [8][8]
    double* sB = sDT + N2;                      // [2][8]
    constexpr int TABI = Grad2Smem::tabInts;
    int* sTab = (int*)(sB + 2 * n);
    uint64_t* bar = (uint64_t*)(((uintptr_t)(sTab + 2 * TABI) + 7) & ~(uintptr_t)7);   // bar[0]: Q, bar[1..2]: face tables
    const size_t es = (size_t)m.nElem * N3;
    const int nTiles = eEnd - eBegin;
    const int tid = threadIdx.x;
    if (tid < N2) sDT[tid] = m.DT[tid];
    if (tid < 2 * n) sB[tid] = m.b[tid];
    auto issueTab = [&](int tile, int buf) {
        const int e0 = eBegin + tile;
        int* dst = sTab + buf * TABI;
        mbar_arrive_expect_tx(bar + 1 + buf, (uint32_t)(TABI * sizeof(int)));
        bulk_g2s(dst, m.elemTrace + (size_t)e0 * 6 * N2, (uint32_t)(6 * N2 * sizeof(int)), bar + 1 + buf);
        bulk_g2s(dst + 6 * N2, m.elemInfo + (size_t)e0 * 8, (uint32_t)(8 * sizeof(int)), bar + 1 + buf);
    };
    auto issueQ = [&](int tile) {
        const size_t off = (size_t)(eBegin + tile) * N3;
        mbar_arrive_expect_tx(bar, 5u * N3 * sizeof(double));
#pragma unroll 1
        for (int c = 0; c < 5; ++c) bulk_g2s(sQ + c * N3, m.Q + c * es + off, N3 * sizeof(double), bar);
    };
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); fence_barrier_init(); }
    __syncthreads();
    auto prefetchPlain = [&](int tile) {   // metrics and 1/J are read with plain loads: into L2 one tile ahead
        const size_t off = (size_t)(eBegin + tile) * N3;
#pragma unroll 1
        for (int c = 0; c < 9; ++c) bulk_prefetch_l2(m.Ja + c * es + off, N3 * sizeof(double));
        bulk_prefetch_l2(m.invJ + off, N3 * sizeof(double));
    };
    if (tid == 0 && (int)blockIdx.x < nTiles) { issueTab(blockIdx.x, 0); issueQ(blockIdx.x); prefetchPlain(blockIdx.x); }
    uint32_t parity = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++iter) {
        const int e = eBegin + tile;
        const int next = tile + (int)gridDim.x;
        const int tbuf = iter & 1;
        const int* sTr = sTab + tbuf * TABI;
        const int* sInfo = sTr + 6 * N2;
        if (tid == 0 && next < nTiles) { issueTab(next, tbuf ^ 1); prefetchPlain(next); }
        mbar_wait(bar + 1 + tbuf, (uint32_t)((iter >> 1) & 1));
        // interface data of the six faces at element-trace nodes
        {
            GradIface gi[IFI];
#pragma unroll
            for (int it = 0; it < IFI; ++it) {
                const int o = tid + it * NT;
                if (o < 6 * N2) {
                    const int lf = o >> 6, ab = o & 63;
                    const size_t fsz = (size_t)m.nFace * N2;
                    gi[it].info = sInfo[lf];
                    const size_t fo = (size_t)sTr[lf * N2 + ab];
                    gi[it].Jf = m.fJ[fo];
#pragma unroll
                    for (int d = 0; d < 3; ++d) gi[it].nh[d] = m.fN[d * fsz + fo];
#pragma unroll
                    for (int q = 0; q < 5; ++q) { gi[it].QL[q] = m.fQ[(size_t)q * fsz + fo]; gi[it].QR[q] = m.fQ[(size_t)(5 + q) * fsz + fo]; }
                }
            }
#pragma unroll
            for (int it = 0; it < IFI; ++it) {
                const int o = tid + it * NT;
                if (o < 6 * N2) {
                    const int lf = o >> 6, ab = o & 63;
                    grad_iface_store<n, false>(m, ph, gi[it], sH + (lf * 5) * N2 + ab, sNrm + (lf * 4) * N2 + ab);
                }
            }
        }
        mbar_wait(bar, parity); parity ^= 1;
        __syncthreads();
        if constexpr (MMA) mma_gradient_contract<NT / 32, N3, N3, true>(sQ, sG, sDT);   // U_xi, U_eta, U_zeta of the 5 variables -> sG
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            const int node = tid + r * NT;
            const int i = node & 7, j = (node >> 3) & 7, k = node >> 6;
            const size_t go = (size_t)e * N3 + node;
            const int p = pswz(node);
            double Uxi[5] = {0, 0, 0, 0, 0}, Ueta[5] = {0, 0, 0, 0, 0}, Uzeta[5] = {0, 0, 0, 0, 0};
            if (MMA) {
#pragma unroll
                for (int q = 0; q < 5; ++q) { Uxi[q] = sG[q * N3 + p]; Ueta[q] = sG[(5 + q) * N3 + p]; Uzeta[q] = sG[(10 + q) * N3 + p]; }
            } else {
                const int bx = (k * n + j) * n, by = (k * n) * n + i, bz = j * n + i;
#pragma unroll
                for (int l = 0; l < n; ++l) {
                    const double dx = sDT[l * n + i], dy = sDT[l * n + j], dz = sDT[l * n + k];
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        Uxi[q] = Uxi[q] + sQ[q * N3 + bx + l] * dx;
                        Ueta[q] = Ueta[q] + sQ[q * N3 + by + l * n] * dy;
                        Uzeta[q] = Uzeta[q] + sQ[q * N3 + bz + l * N2] * dz;
                    }
                }
            }
            asm volatile("" ::: "memory");   // keep the metric loads below the contraction: hoisted, they cost 600 bytes of spills
            double ja[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) ja[c] = m.Ja[c * es + go];
            const double iJ = m.invJ[go];
            double g[15];
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
                const double* H = sH + (lf * 5) * N2 + ab;
                const double* Nn = sNrm + (lf * 4) * N2 + ab;
                const bool bnd = ((sInfo[lf] >> 4) & 3) == H3D_FACE_BOUNDARY;
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
            double* ox = m.Ux + go; double* oy = m.Uy + go; double* oz = m.Uz + go;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                // Euler with "compute gradients": local gradient only (base-class ComputeGradient, EllipticDiscretizationClass.f90:122-187)
                if (ph.ns) { g[q] = g[q] + fx[q] * iJ; g[5 + q] = g[5 + q] + fy[q] * iJ; g[10 + q] = g[10 + q] + fz[q] * iJ; }
                ox[q * es] = g[q]; oy[q * es] = g[5 + q]; oz[q * es] = g[10 + q];
            }
#pragma unroll
            for (int c = 0; c < 15; ++c) sG[c * N3 + p] = g[c];
        }
        __syncthreads();   // Q and the interface data are consumed, the gradients are complete
        if (tid == 0 && next < nTiles) issueQ(next);
        prolong2_block<15>(m, ops, sG, sTr, sInfo, m.fU);
        __syncthreads();
    }
}

}  // namespace h3d
