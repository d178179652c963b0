// Host-side geometry of p-nonconforming meshes: every element has its own polynomial orders (Nx, Ny, Nz), every face the
// orders of its two sides and its own (the maximum per direction).  Driver-side pre-processing like geometry.hpp; the device
// library receives the flat arrays (include/h3d_gpu.h, h3d_set_mesh_p).
//
// Reference behaviour followed (paths relative to /root/reference/Solver/src/libs):
//   mesh/FaceClass.f90:187-282                   Face_LinkWithElements: NelLeft/Right, NfLeft/Right, Nf = max, projectionType
//   mesh/HexMesh.f90:2399-2480                   element face orders through axisMap; boundary faces: NelRight = NelLeft
//   mesh/HexMesh.f90:2087-2320                   HexMesh_CheckIfMeshIs2D (extruded meshes skip the boundary-order rule below)
//   mesh/HexMesh.f90:2730-2795, 5963-5998        boundary mapping order of anisotropic 3D meshes: min order of the zone, halved
//                                                unless the representation is conforming on the zone
//   mesh/HexMesh.f90:2797-2960                   curved patches re-sampled on the Chebyshev-Lobatto points of the lower order
//                                                (ProjectFaceToNewPoints, FacePatchClass.f90:278-300) so that both sides of a
//                                                face see the same surface
//   mesh/MappedGeometry.f90:64-152, 174-384      element geometry with three nodal storages (anisotropic)
//   mesh/MappedGeometry.f90:482-756              face geometry from the LEFT element at its order, interpolated to the face
//                                                order with Tset (spectral/InterpolationMatrices.f90:42-107)
#pragma once
#include <map>

#include "geometry.hpp"

namespace h3d {

// InterpolationMatrix_Construct (InterpolationMatrices.f90:42-107): T is (Ndest+1) x (Norigin+1) row-major, T[i*(Norigin+1)+l] = T(i,l).
// Norigin < Ndest: Lagrange interpolation; otherwise the transposed interpolation weighted by w_origin(j) / w_dest(i) (L2 projection).
inline void interpolationMatrix(const NodalStorage& spAo, const NodalStorage& spAd, std::vector<double>& T) {
    const int No = spAo.N, Nd = spAd.N;
    T.assign((size_t)(Nd + 1) * (No + 1), 0.0);
    if (No < Nd) {
        polynomialInterpolationMatrix(No, Nd, spAo.x.data(), spAo.wb.data(), spAd.x.data(), T.data());
    } else {
        std::vector<double> Tt((size_t)(No + 1) * (Nd + 1));
        polynomialInterpolationMatrix(Nd, No, spAd.x.data(), spAd.wb.data(), spAo.x.data(), Tt.data());
        for (int j = 0; j <= No; ++j) for (int i = 0; i <= Nd; ++i) T[(size_t)i * (No + 1) + j] = Tt[(size_t)j * (Nd + 1) + i] * spAo.w[j] / spAd.w[i];
    }
}

struct HostGeometryP {
    int nodeType = GAUSS;
    bool meshIs2D = false, anisotropic = false;
    std::vector<int> elemOrder;                 // [e][3]  Nx, Ny, Nz
    std::vector<int> faceOrder;                 // [f][6]  Nf(1:2), NfLeft(1:2), NfRight(1:2)
    std::vector<long long> eOff, fOff;          // node offsets, nElem+1 / nFaces+1
    std::map<int, NodalStorage> sp;             // NodalStorage(N)
    // element arrays: elements concatenated, [k][j][i][c] inside an element (the reference's packed order)
    std::vector<double> x, jGradXi, jGradEta, jGradZeta, jac, invJac, volume;
    // face arrays at the face order: [j][i][c]
    std::vector<double> fx, fnormal, ft1, ft2, fjac, fsurface;
    std::vector<double> fh;                     // f % geom % h (HexMesh.f90:3016-3041): min(J) of the adjacent elements / max(J_f)
    std::vector<double> dWall, fdWall;          // distance of every element / face node to the nearest no-slip wall node (optional)
    const NodalStorage& S(int N) { auto it = sp.find(N); if (it == sp.end()) { sp[N].construct(nodeType, N); return sp[N]; } return it->second; }
};

// HexMesh_CheckIfMeshIs2D (HexMesh.f90:2087-2320): every element has a local direction whose four edges are parallel to the
// same global axis
inline bool meshIsExtruded(const HostMesh& m) {
    const int pairs[3][2] = {{5, 3}, {0, 1}, {2, 4}};   // (ELEFT, ERIGHT), (EFRONT, EBACK), (EBOTTOM, ETOP)
    int oriented[3] = {0, 0, 0};
    for (int e = 0; e < m.nElem(); ++e)
        for (int p = 0; p < 3; ++p) for (int dir = 0; dir < 3; ++dir) {
            int cnt = 0;
            for (int q = 0; q < 4; ++q) {
                const double* x1 = &m.nodes[3 * m.elemNodes[8 * e + localFaceNode[pairs[p][0]][q]]];
                const double* x2 = &m.nodes[3 * m.elemNodes[8 * e + localFaceNode[pairs[p][1]][q]]];
                double dx[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
                const double nrm = std::sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
                if (almostEqual(std::fabs(dx[dir] / nrm), 1.0)) ++cnt;
            }
            if (cnt == 4) ++oriented[dir];
        }
    return oriented[0] == m.nElem() || oriented[1] == m.nElem() || oriented[2] == m.nElem();
}

inline bool buildGeometryP(const HostMesh& m0, const int* Nxyz, int nodeType, HostGeometryP& g, std::string& err) {
    const int nE = m0.nElem(), nF = m0.nFaces;
    g = HostGeometryP(); g.nodeType = nodeType;
    g.elemOrder.assign(Nxyz, Nxyz + 3 * (size_t)nE);
    int mn = 1 << 30, mx = 0;
    for (int q = 0; q < 3 * nE; ++q) { mn = std::min(mn, Nxyz[q]); mx = std::max(mx, Nxyz[q]); }
    if (mn < 1 || mx > 15) { err = "polynomial orders must lie in 1..15"; return false; }
    g.anisotropic = mn != mx;                               // DGSEMClass.f90:213
    g.meshIs2D = meshIsExtruded(m0);
    for (int f = 0; f < nF; ++f) if (m0.faceType[f] == HMESH_MPI) { err = "p-nonconforming meshes are single-domain (no MPI faces)"; return false; }
    // ---- Face_LinkWithElements
    g.faceOrder.assign(6 * (size_t)nF, 0);
    std::vector<int> projType(2 * (size_t)nF, 0);
    for (int f = 0; f < nF; ++f) {
        const int eL = m0.faceElem[2 * f], lfL = m0.faceElemSide[2 * f], eR = m0.faceElem[2 * f + 1], lfR = m0.faceElemSide[2 * f + 1];
        int NelL[2] = {Nxyz[3 * eL + axisMap[lfL][0]], Nxyz[3 * eL + axisMap[lfL][1]]}, NelR[2] = {NelL[0], NelL[1]};
        if (m0.faceType[f] == HMESH_INTERIOR) { NelR[0] = Nxyz[3 * eR + axisMap[lfR][0]]; NelR[1] = Nxyz[3 * eR + axisMap[lfR][1]]; }
        int NfR[2] = {NelR[0], NelR[1]};
        const int rot = m0.faceRot[f];
        if (rot == 1 || rot == 3 || rot == 4 || rot == 6) { NfR[0] = NelR[1]; NfR[1] = NelR[0]; }
        int* fo = &g.faceOrder[6 * (size_t)f];
        fo[0] = std::max(NelL[0], NfR[0]); fo[1] = std::max(NelL[1], NfR[1]); fo[2] = NelL[0]; fo[3] = NelL[1]; fo[4] = NfR[0]; fo[5] = NfR[1];
        projType[2 * f] = (fo[2] != fo[0] ? 1 : 0) + (fo[3] != fo[1] ? 2 : 0);
        projType[2 * f + 1] = (fo[4] != fo[0] ? 1 : 0) + (fo[5] != fo[1] ? 2 : 0);
    }
    g.eOff.assign(nE + 1, 0); g.fOff.assign(nF + 1, 0);
    for (int e = 0; e < nE; ++e) g.eOff[e + 1] = g.eOff[e] + (long long)(Nxyz[3 * e] + 1) * (Nxyz[3 * e + 1] + 1) * (Nxyz[3 * e + 2] + 1);
    for (int f = 0; f < nF; ++f) g.fOff[f + 1] = g.fOff[f] + (long long)(g.faceOrder[6 * f] + 1) * (g.faceOrder[6 * f + 1] + 1);
    for (int N = 1; N <= mx; ++N) g.S(N);
    // ---- boundary mapping orders of anisotropic 3D meshes (HexMesh.f90:2742-2795)
    const int nZ = (int)m0.bcs.size();
    std::vector<int> bfOrder(nZ, 1 << 30);
    const bool zoneRule = g.anisotropic && !g.meshIs2D;
    if (zoneRule) {
        static const int neighborFaces[6][4] = {{2, 3, 4, 5}, {2, 3, 4, 5}, {0, 1, 3, 5}, {0, 1, 2, 4}, {0, 1, 3, 5}, {0, 1, 2, 4}};   // HexElementConnectivityDefinitions.f90
        std::vector<char> conforming(nZ, 1), used(nZ, 0);
        for (int f = 0; f < nF; ++f) {
            if (m0.faceType[f] != HMESH_BOUNDARY) continue;
            const int z = m0.faceZone[f]; used[z] = 1;
            bfOrder[z] = std::min(bfOrder[z], std::min(g.faceOrder[6 * f + 2], g.faceOrder[6 * f + 3]));
            const int e = m0.faceElem[2 * f], lf = m0.faceElemSide[2 * f];
            for (int q = 0; q < 4; ++q) {   // HexMesh_ConformingOnZone
                const int nf = m0.elemFace[6 * e + neighborFaces[lf][q]];
                if (m0.faceType[nf] == HMESH_BOUNDARY) continue;
                if (g.faceOrder[6 * nf + 2] != g.faceOrder[6 * nf + 4] || g.faceOrder[6 * nf + 3] != g.faceOrder[6 * nf + 5]) conforming[z] = 0;
            }
        }
        for (int z = 0; z < nZ; ++z) if (used[z] && !conforming[z]) {
            bfOrder[z] = bfOrder[z] / 2;
            if (bfOrder[z] < 1) { err = "The chosen polynomial orders are too low to represent the boundaries accurately (nonconforming representations on boundaries need N>=2)"; return false; }
        }
    }
    // ---- local copy of the surface patches, adapted to the solution order (HexMesh.f90:2797-2960)
    HostMesh m = m0;
    auto project = [&](int e, int lf, const int CLN[2]) {
        ElemMap map; map.init(m, e);
        FacePatch np; np.nu = CLN[0] + 1; np.nv = CLN[1] + 1; np.pts.resize(3 * (size_t)np.nu * np.nv);
        const NodalStorage &s1 = g.S(CLN[0]), &s2 = g.S(CLN[1]);
        for (int j = 0; j < np.nv; ++j) for (int i = 0; i < np.nu; ++i) map.facePoint(lf, s1.xCGL[i], s2.xCGL[j], &np.pts[3 * ((size_t)j * np.nu + i)]);
        m.patches[e][lf] = np;
    };
    auto patchOrder = [&](int e, int lf, int NS[2]) { NS[0] = m.patches[e][lf].nu - 1; NS[1] = m.patches[e][lf].nv - 1; };
    for (int f = 0; f < nF; ++f) {
        const int eL = m.faceElem[2 * f], lfL = m.faceElemSide[2 * f];
        const int* fo = &g.faceOrder[6 * (size_t)f];
        if (m.faceType[f] == HMESH_INTERIOR) {
            const int eR = m.faceElem[2 * f + 1], lfR = m.faceElemSide[2 * f + 1];
            int NSL[2] = {1, 1}, NSR[2] = {1, 1};
            if (!m.isHex8[eL]) patchOrder(eL, lfL, NSL);
            if (!m.isHex8[eR]) patchOrder(eR, lfR, NSR);
            if (m.isHex8[eL] && m.isHex8[eR]) continue;
            if (m.isHex8[eL] && NSR[0] == 1 && NSR[1] == 1) continue;
            if (m.isHex8[eR] && NSL[0] == 1 && NSL[1] == 1) continue;
            if (NSL[0] == 1 && NSL[1] == 1 && NSR[0] == 1 && NSR[1] == 1) continue;
            int CLN[2] = {std::min(fo[2], fo[4]), std::min(fo[3], fo[5])};
            if (!m.isHex8[eL] && (CLN[0] < NSL[0] || CLN[1] < NSL[1])) project(eL, lfL, CLN);
            const int rot = m.faceRot[f];
            if ((rot == 1 || rot == 3 || rot == 4 || rot == 6) && CLN[0] != CLN[1]) std::swap(CLN[0], CLN[1]);
            if (!m.isHex8[eR] && (CLN[0] < NSR[0] || CLN[1] < NSR[1])) project(eR, lfR, CLN);
        } else {
            if (m.isHex8[eL]) continue;
            int NSL[2]; patchOrder(eL, lfL, NSL);
            if (NSL[0] == 1 && NSL[1] == 1) continue;
            int CLN[2] = {fo[2], fo[3]};
            if (zoneRule) CLN[0] = CLN[1] = bfOrder[m.faceZone[f]];
            if (CLN[0] < NSL[0] || CLN[1] < NSL[1]) project(eL, lfL, CLN);
        }
    }
    // ---- elements (ConstructMappedGeometry + computeMetricTermsConservativeForm)
    const size_t nn = (size_t)g.eOff[nE];
    g.x.assign(3 * nn, 0.0); g.jGradXi.assign(3 * nn, 0.0); g.jGradEta.assign(3 * nn, 0.0); g.jGradZeta.assign(3 * nn, 0.0);
    g.jac.assign(nn, 0.0); g.invJac.assign(nn, 0.0); g.volume.assign(nE, 0.0);
#pragma omp parallel
    {
        std::vector<double> xC, gradx, cp, aux, Ja[3], JC;
        ElemMap map;
#pragma omp for schedule(dynamic, 4)
        for (int e = 0; e < nE; ++e) {
            const int Nx = Nxyz[3 * e], Ny = Nxyz[3 * e + 1], Nz = Nxyz[3 * e + 2], nx = Nx + 1, ny = Ny + 1, nz = Nz + 1, n3 = nx * ny * nz;
            const NodalStorage &sx = g.sp.at(Nx), &sy = g.sp.at(Ny), &sz = g.sp.at(Nz);
            const size_t o = (size_t)g.eOff[e];
            map.init(m, e);
            auto I = [&](int i, int j, int k) { return (k * ny + j) * nx + i; };
            xC.assign(3 * n3, 0.0); gradx.assign(9 * n3, 0.0); cp.assign(3 * n3, 0.0); aux.assign(9 * n3, 0.0); JC.assign(n3, 0.0);
            for (int d = 0; d < 3; ++d) Ja[d].assign(3 * n3, 0.0);
            for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
                double u[3] = {sx.x[i], sy.x[j], sz.x[k]};
                map.at(u, &g.x[3 * (o + I(i, j, k))]);
                double uc[3] = {sx.xCGL[i], sy.xCGL[j], sz.xCGL[k]}, gg[3][3];
                map.at(uc, &xC[3 * I(i, j, k)]);
                if (m.isHex8[e]) map.gradHex8(uc, gg); else map.gradGeneral(uc, gg);
                for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) gradx[9 * I(i, j, k) + 3 * a + b] = gg[a][b];
            }
            for (int c = 0; c < 3; ++c) {
                const int l_ = (c + 2) % 3, m_ = (c + 1) % 3;
                for (int q = 0; q < n3; ++q) for (int d = 0; d < 3; ++d)
                    cp[3 * q + d] = xC[3 * q + l_] * gradx[9 * q + 3 * m_ + d] - xC[3 * q + m_] * gradx[9 * q + 3 * l_ + d];
                std::fill(aux.begin(), aux.end(), 0.0);
                for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
                    double* a = &aux[9 * I(i, j, k)];   // a[3*comp + dir]
                    for (int l = 0; l < nx; ++l) for (int d = 0; d < 3; ++d) a[3 * d + 0] += cp[3 * I(l, j, k) + d] * sx.DCGL[i * nx + l];
                    for (int l = 0; l < ny; ++l) for (int d = 0; d < 3; ++d) a[3 * d + 1] += cp[3 * I(i, l, k) + d] * sy.DCGL[j * ny + l];
                    for (int l = 0; l < nz; ++l) for (int d = 0; d < 3; ++d) a[3 * d + 2] += cp[3 * I(i, j, l) + d] * sz.DCGL[k * nz + l];
                }
                for (int q = 0; q < n3; ++q) {
                    const double* a = &aux[9 * q];
                    const double J1 = a[3 * 2 + 1] - a[3 * 1 + 2], J2 = a[3 * 0 + 2] - a[3 * 2 + 0], J3 = a[3 * 1 + 0] - a[3 * 0 + 1];
                    Ja[0][3 * q + c] = -0.5 * J1; Ja[1][3 * q + c] = -0.5 * J2; Ja[2][3 * q + c] = -0.5 * J3;
                }
            }
            for (int q = 0; q < n3; ++q) {
                const double* G = &gradx[9 * q];
                const double a1[3] = {G[0], G[3], G[6]}, a2[3] = {G[1], G[4], G[7]}, a3[3] = {G[2], G[5], G[8]};
                JC[q] = a1[0] * (a2[1] * a3[2] - a2[2] * a3[1]) + a1[1] * (a2[2] * a3[0] - a2[0] * a3[2]) + a1[2] * (a2[0] * a3[1] - a2[1] * a3[0]);
            }
            // back to the solution nodes: the reference's triple sum (MappedGeometry.f90:336-361)
            auto interp3 = [&](const double* src, int nc, double* dst) {
                for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
                    double acc[3] = {0.0, 0.0, 0.0};
                    for (int nn_ = 0; nn_ < nz; ++nn_) for (int mm = 0; mm < ny; ++mm) for (int l = 0; l < nx; ++l)
                        for (int c = 0; c < nc; ++c)
                            acc[c] = acc[c] + src[nc * I(l, mm, nn_) + c] * sx.TCheb2Gauss[i * nx + l] * sy.TCheb2Gauss[j * ny + mm] * sz.TCheb2Gauss[k * nz + nn_];
                    for (int c = 0; c < nc; ++c) dst[nc * I(i, j, k) + c] = acc[c];
                }
            };
            interp3(Ja[0].data(), 3, &g.jGradXi[3 * o]); interp3(Ja[1].data(), 3, &g.jGradEta[3 * o]); interp3(Ja[2].data(), 3, &g.jGradZeta[3 * o]);
            interp3(JC.data(), 1, &g.jac[o]);
            double vol = 0.0;
            for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
                const size_t q = o + I(i, j, k);
                g.invJac[q] = 1.0 / g.jac[q];
                vol = vol + sx.w[i] * sy.w[j] * sz.w[k] * g.jac[q];
            }
            g.volume[e] = vol;
        }
    }
    // ---- faces (ConstructMappedGeometryFace, from the LEFT element: side = 1, rot = 0)
    const size_t nfn = (size_t)g.fOff[nF];
    g.fx.assign(3 * nfn, 0.0); g.fnormal.assign(3 * nfn, 0.0); g.ft1.assign(3 * nfn, 0.0); g.ft2.assign(3 * nfn, 0.0); g.fjac.assign(nfn, 0.0); g.fsurface.assign(nF, 0.0);
    std::map<std::pair<int, int>, std::vector<double>> Tset;
    for (int f = 0; f < nF; ++f) for (int d = 0; d < 2; ++d) {
        const std::pair<int, int> key(g.faceOrder[6 * f + 2 + d], g.faceOrder[6 * f + d]);
        if (!Tset.count(key)) interpolationMatrix(g.sp.at(key.first), g.sp.at(key.second), Tset[key]);
    }
#pragma omp parallel
    {
        ElemMap map;
        std::vector<double> dS, nrmF;
#pragma omp for schedule(dynamic, 8)
        for (int f = 0; f < nF; ++f) {
            const int e = m.faceElem[2 * f], lf = m.faceElemSide[2 * f];
            const int* fo = &g.faceOrder[6 * (size_t)f];
            const int Nf1 = fo[0], Nf2 = fo[1], Ne1 = fo[2], Ne2 = fo[3], nf1 = Nf1 + 1, nf2 = Nf2 + 1, ne1 = Ne1 + 1, ne2 = Ne2 + 1;
            const int nx = Nxyz[3 * e] + 1, ny = Nxyz[3 * e + 1] + 1, nz = Nxyz[3 * e + 2] + 1;
            const int nrmAxis = faceNormalAxis[lf], nN = (nrmAxis == 0 ? nx : nrmAxis == 1 ? ny : nz);
            const NodalStorage &s1 = g.sp.at(Nf1), &s2 = g.sp.at(Nf2), &sn = g.sp.at(nN - 1);
            const size_t fo0 = (size_t)g.fOff[f], eo = (size_t)g.eOff[e];
            map.init(m, e);
            for (int j = 0; j < nf2; ++j) for (int i = 0; i < nf1; ++i) {
                double u[3]; u[axisMap[lf][0]] = s1.x[i]; u[axisMap[lf][1]] = s2.x[j]; u[nrmAxis] = faceNormalEnd[lf] ? 1.0 : -1.0;
                map.at(u, &g.fx[3 * (fo0 + (size_t)j * nf1 + i)]);
            }
            const double* v = &sn.v[faceNormalEnd[lf] * nN];
            const double* Jd = (nrmAxis == 0 ? g.jGradXi.data() : nrmAxis == 1 ? g.jGradEta.data() : g.jGradZeta.data()) + 3 * eo;
            dS.assign(3 * (size_t)ne1 * ne2, 0.0);
            for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
                const int idx[3] = {i, j, k};
                const int a = idx[axisMap[lf][0]], b = idx[axisMap[lf][1]], nr = idx[nrmAxis];
                for (int c = 0; c < 3; ++c) dS[3 * (b * ne1 + a) + c] = dS[3 * (b * ne1 + a) + c] + Jd[3 * ((k * ny + j) * nx + i) + c] * v[nr];
            }
            if (!faceNormalEnd[lf]) for (auto& t : dS) t = -t;
            nrmF.assign(3 * (size_t)nf1 * nf2, 0.0);
            const int pt = projType[2 * f];
            const double* T1 = Tset.at({Ne1, Nf1}).data(); const double* T2 = Tset.at({Ne2, Nf2}).data();
            if (pt == 0) nrmF = dS;
            else if (pt == 1) {
                for (int j = 0; j < nf2; ++j) for (int l = 0; l < ne1; ++l) for (int i = 0; i < nf1; ++i) for (int c = 0; c < 3; ++c)
                    nrmF[3 * (j * nf1 + i) + c] = nrmF[3 * (j * nf1 + i) + c] + T1[i * ne1 + l] * dS[3 * (j * ne1 + l) + c];
            } else if (pt == 2) {
                for (int l = 0; l < ne2; ++l) for (int j = 0; j < nf2; ++j) for (int i = 0; i < nf1; ++i) for (int c = 0; c < 3; ++c)
                    nrmF[3 * (j * nf1 + i) + c] = nrmF[3 * (j * nf1 + i) + c] + T2[j * ne2 + l] * dS[3 * (l * ne1 + i) + c];
            } else {
                for (int l = 0; l < ne2; ++l) for (int j = 0; j < nf2; ++j) for (int mm = 0; mm < ne1; ++mm) for (int i = 0; i < nf1; ++i) for (int c = 0; c < 3; ++c)
                    nrmF[3 * (j * nf1 + i) + c] = nrmF[3 * (j * nf1 + i) + c] + T1[i * ne1 + mm] * T2[j * ne2 + l] * dS[3 * (l * ne1 + mm) + c];
            }
            double surf = 0.0;
            for (int j = 0; j < nf2; ++j) for (int i = 0; i < nf1; ++i) {
                const size_t q = fo0 + (size_t)j * nf1 + i;
                const double* nv = &nrmF[3 * (j * nf1 + i)];
                const double nrm = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
                g.fjac[q] = nrm;
                for (int c = 0; c < 3; ++c) g.fnormal[3 * q + c] = nv[c] / nrm;
            }
            for (int j = 0; j < nf2; ++j) for (int i = 0; i < nf1; ++i) {
                const size_t q = fo0 + (size_t)j * nf1 + i;
                double t1[3] = {0, 0, 0};
                for (int l = 0; l < nf1; ++l) for (int c = 0; c < 3; ++c) t1[c] += s1.D[i * nf1 + l] * g.fx[3 * (fo0 + (size_t)j * nf1 + l) + c];
                const double* nh = &g.fnormal[3 * q];
                const double dot = t1[0] * nh[0] + t1[1] * nh[1] + t1[2] * nh[2];
                for (int c = 0; c < 3; ++c) t1[c] -= dot * nh[c];
                const double nt = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
                for (int c = 0; c < 3; ++c) { t1[c] /= nt; g.ft1[3 * q + c] = t1[c]; }
                g.ft2[3 * q + 0] = nh[1] * t1[2] - nh[2] * t1[1];
                g.ft2[3 * q + 1] = nh[2] * t1[0] - nh[0] * t1[2];
                g.ft2[3 * q + 2] = nh[0] * t1[1] - nh[1] * t1[0];
                surf = surf + s1.w[i] * s2.w[j] * g.fjac[q];
            }
            g.fsurface[f] = surf;
        }
    }
    // faces' minimum orthogonal distance estimate (HexMesh.f90:3016-3041)
    std::vector<double> minJ(nE);
    for (int e = 0; e < nE; ++e) minJ[e] = *std::min_element(g.jac.begin() + g.eOff[e], g.jac.begin() + g.eOff[e + 1]);
    g.fh.assign(nF, 0.0);
    for (int f = 0; f < nF; ++f) {
        const int e1 = m.faceElem[2 * f], e2 = m.faceElem[2 * f + 1];
        const double num = (e1 >= 0 && e2 >= 0) ? std::min(minJ[e1], minJ[e2]) : minJ[std::max(e1, e2)];
        g.fh[f] = num / *std::max_element(g.fjac.begin() + g.fOff[f], g.fjac.begin() + g.fOff[f + 1]);
    }
    return true;
}

// HexMesh_ComputeWallDistances (HexMesh.f90:5594-5692) on a p-nonconforming mesh: the wall nodes are the nodes of the no-slip faces at
// their face orders
inline std::vector<double> wallCoordinatesP(const HostMesh& m, const HostGeometryP& g) {
    std::vector<double> Xw;
    for (int f = 0; f < m.nFaces; ++f) {
        if (m.faceType[f] != HMESH_BOUNDARY || m.faceZone[f] < 0 || m.bcs[m.faceZone[f]].type != "noslipwall") continue;
        Xw.insert(Xw.end(), &g.fx[3 * (size_t)g.fOff[f]], &g.fx[3 * (size_t)g.fOff[f + 1]]);
    }
    return Xw;
}
inline void computeWallDistancesP(HostGeometryP& g, const std::vector<double>& Xw) {
    const size_t nW = Xw.size() / 3;
    auto dist = [&](const double* xP) {
        double mn = 1.7976931348623157e308;
        for (size_t q = 0; q < nW; ++q) {
            const double d0 = xP[0] - Xw[3 * q], d1 = xP[1] - Xw[3 * q + 1], d2 = xP[2] - Xw[3 * q + 2];
            mn = std::fmin(mn, d0 * d0 + d1 * d1 + d2 * d2);
        }
        return std::sqrt(mn);
    };
    g.dWall.resize(g.x.size() / 3); g.fdWall.resize(g.fx.size() / 3);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)g.dWall.size(); ++q) g.dWall[q] = dist(&g.x[3 * q]);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)g.fdWall.size(); ++q) g.fdWall[q] = dist(&g.fx[3 * q]);
}

}  // namespace h3d
