// Host-side hexahedral mesh: SpecMesh reader, synthetic periodic-box generator, face connectivity,
// periodic pairing.  Plays the role of the reference's HexMesh construction on the driver side; the
// device library only ever sees the flat arrays produced here (include/h3d_gpu.h, h3d_set_mesh).
//
// Reference behaviour followed (paths relative to /root/reference/Solver/src/libs/mesh):
//   Read_SpecMesh.f90:116,133-136,176-245      file layout, Chebyshev-Lobatto face patches, flat patches
//   HexElementConnectivityDefinitions.f90:36-77 axisMap, localFaceNode
//   HexMesh.f90:356-432                        ConstructFaces (first element to touch a face is LEFT)
//   HexMesh.f90:490-515                        faceRotation
//   HexMesh.f90:519-736, 740-804               ConstructPeriodicFaces / CompareTwoNodes
//   HexMesh.f90:872-944                        DeletePeriodicMinusFaces
//   HexMesh.f90:434-476                        GetElementsFaceIDs
//   MeshTypes.f90:70-108                       leftIndexes2Right
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "nodal.hpp"

namespace h3d {

enum { HMESH_INTERIOR = 1, HMESH_BOUNDARY = 2, HMESH_MPI = 3, HMESH_UNDEFINED = 0 };

// 0-based local corner ids of the 6 local faces (reference numbering minus one)
static const int localFaceNode[6][4] = {{0, 1, 5, 4}, {3, 2, 6, 7}, {0, 1, 2, 3}, {1, 2, 6, 5}, {4, 5, 6, 7}, {0, 3, 7, 4}};
// element axes spanned by the face (0=xi,1=eta,2=zeta)
static const int axisMap[6][2] = {{0, 2}, {0, 2}, {0, 1}, {1, 2}, {0, 1}, {1, 2}};
// normal axis and end (0: -1 side, 1: +1 side)
static const int faceNormalAxis[6] = {1, 1, 2, 0, 2, 0};
static const int faceNormalEnd[6] = {0, 1, 0, 1, 1, 0};

inline void leftIndexes2Right(int i, int j, int Nx, int Ny, int rot, int& ii, int& jj) {
    switch (rot) {
        case 0: ii = i; jj = j; break;
        case 1: ii = Ny - j; jj = i; break;
        case 2: ii = Nx - i; jj = Ny - j; break;
        case 3: ii = j; jj = Nx - i; break;
        case 4: ii = j; jj = i; break;
        case 5: ii = Nx - i; jj = j; break;
        case 6: ii = Ny - j; jj = Nx - i; break;
        case 7: ii = i; jj = Ny - j; break;
        default: ii = i; jj = j;
    }
}

inline int faceRotation(const int* master, const int* slave) {
    static const int NEXT[4] = {1, 2, 3, 0};
    int j = 0;
    for (j = 0; j < 4; ++j) if (master[0] == slave[j]) break;
    if (j == 4) return -1;
    return (master[1] == slave[NEXT[j]]) ? j : j + 4;
}

struct FacePatch {            // points[(j*nu + i)*3 + c], knots are Chebyshev-Lobatto of order nu-1 / nv-1
    int nu = 0, nv = 0;
    std::vector<double> pts;
};

struct BCSpec {
    std::string name;         // lower case
    std::string type;         // "periodic", "noslipwall", "freeslipwall", "inflow", "outflow"
    std::string coupled;      // periodic partner
    double params[16] = {0};
};

struct HostMesh {
    // --- raw mesh
    int bFaceOrder = 1;
    std::vector<double> nodes;                 // 3*nNodes
    std::vector<int> elemNodes;                // 8*nElem (0-based)
    std::vector<char> isHex8;                  // nElem
    std::vector<std::array<FacePatch, 6>> patches;  // only meaningful when !isHex8
    std::vector<std::string> bname;            // 6*nElem, lower case, "---" = interior
    // --- connectivity
    std::vector<BCSpec> bcs;
    int nFaces = 0;
    std::vector<int> faceNodes;                // 4*nFaces
    std::vector<int> faceElem;                 // 2*nFaces  (-1 = none)
    std::vector<int> faceElemSide;             // 2*nFaces  local face id 0..5 (-1 = none)
    std::vector<int> faceRot, faceType, faceZone;
    std::vector<int> elemFace, elemFaceSide;   // 6*nElem ; side 0 = left, 1 = right
    int nElem() const { return (int)(elemNodes.size() / 8); }
    int nNodes() const { return (int)(nodes.size() / 3); }
};

inline std::string toLower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

// ----------------------------------------------------------------------------------------------------
// SpecMesh reader (Read_SpecMesh.f90:116-245).  Lref division is the caller's business (Lref = 1 here).
inline bool readSpecMesh(const std::string& path, HostMesh& m, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "Error opening file: " + path; return false; }
    // Fortran list-directed READ semantics: a READ of `count` items consumes as many records (lines) as it needs and
    // discards the rest of the last record it touched.
    auto readTokens = [&](int count, std::vector<std::string>& out) -> bool {
        out.clear();
        std::string line;
        while ((int)out.size() < count) {
            if (!std::getline(in, line)) return false;
            std::istringstream ss(line);
            std::string tok;
            while ((int)out.size() < count && (ss >> tok)) out.push_back(tok);
        }
        return true;
    };
    std::vector<std::string> tk;
    auto toD = [](const std::string& t) { std::string u = t; for (auto& c : u) if (c == 'd' || c == 'D') c = 'e'; return std::strtod(u.c_str(), nullptr); };
    if (!readTokens(3, tk)) { err = "bad SpecMesh header"; return false; }
    const int nNodes = std::atoi(tk[0].c_str()), nElems = std::atoi(tk[1].c_str()), bfo = std::atoi(tk[2].c_str());
    if (nNodes <= 0 || nElems <= 0 || bfo < 1) { err = "bad SpecMesh header"; return false; }
    m.bFaceOrder = bfo;
    const int nb = bfo + 1;
    m.nodes.resize(3 * (size_t)nNodes);
    for (int j = 0; j < nNodes; ++j) {
        if (!readTokens(3, tk)) { err = "bad node line"; return false; }
        for (int c = 0; c < 3; ++c) m.nodes[3 * j + c] = toD(tk[c]);
    }
    m.elemNodes.resize(8 * (size_t)nElems); m.isHex8.assign(nElems, 1); m.patches.resize(nElems); m.bname.resize(6 * (size_t)nElems);
    for (int l = 0; l < nElems; ++l) {
        if (!readTokens(8, tk)) { err = "unexpected EOF (element nodes)"; return false; }
        for (int k = 0; k < 8; ++k) {
            const int id = std::atoi(tk[k].c_str());
            if (id < 1 || id > nNodes) { err = "bad element node ids"; return false; }
            m.elemNodes[8 * l + k] = id - 1;
        }
        int flags[6];
        if (!readTokens(6, tk)) { err = "unexpected EOF (face flags)"; return false; }
        for (int k = 0; k < 6; ++k) flags[k] = std::atoi(tk[k].c_str());
        int mx = 0; for (int k = 0; k < 6; ++k) mx = std::max(mx, flags[k]);
        if (mx != 0) {
            m.isHex8[l] = 0;
            for (int k = 0; k < 6; ++k) {
                FacePatch& p = m.patches[l][k];
                if (flags[k] == 0) {
                    p.nu = p.nv = 2; p.pts.resize(12);
                    const int* nm = localFaceNode[k];
                    const int order[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
                    for (int q = 0; q < 4; ++q) for (int c = 0; c < 3; ++c)
                        p.pts[(order[q][1] * 2 + order[q][0]) * 3 + c] = m.nodes[3 * m.elemNodes[8 * l + nm[q]] + c];
                } else {
                    p.nu = p.nv = nb; p.pts.resize(3 * (size_t)nb * nb);
                    for (int j = 0; j < nb; ++j) for (int i = 0; i < nb; ++i) {
                        if (!readTokens(3, tk)) { err = "unexpected EOF (patch)"; return false; }
                        for (int c = 0; c < 3; ++c) p.pts[(j * nb + i) * 3 + c] = toD(tk[c]);
                    }
                }
            }
        }
        if (!readTokens(6, tk)) { err = "unexpected EOF (boundary names)"; return false; }
        for (int k = 0; k < 6; ++k) m.bname[6 * l + k] = toLower(tk[k]);
    }
    return true;
}

inline void writeSpecMesh(const std::string& path, const HostMesh& m) {
    FILE* f = std::fopen(path.c_str(), "w");
    std::fprintf(f, " %11d %11d %11d\n", m.nNodes(), m.nElem(), m.bFaceOrder);
    for (int j = 0; j < m.nNodes(); ++j) std::fprintf(f, " %.17g %.17g %.17g\n", m.nodes[3 * j], m.nodes[3 * j + 1], m.nodes[3 * j + 2]);
    for (int l = 0; l < m.nElem(); ++l) {
        for (int k = 0; k < 8; ++k) std::fprintf(f, " %11d", m.elemNodes[8 * l + k] + 1);
        std::fprintf(f, "\n");
        for (int k = 0; k < 6; ++k) std::fprintf(f, " %11d", m.isHex8[l] ? 0 : 1);
        std::fprintf(f, "\n");
        if (!m.isHex8[l]) for (int k = 0; k < 6; ++k) {
            const FacePatch& p = m.patches[l][k];
            for (int q = 0; q < p.nu * p.nv; ++q) std::fprintf(f, " %.17g %.17g %.17g\n", p.pts[3 * q], p.pts[3 * q + 1], p.pts[3 * q + 2]);
        }
        for (int k = 0; k < 6; ++k) std::fprintf(f, " %-15s", m.bname[6 * l + k].c_str());
        std::fprintf(f, "\n");
    }
    std::fclose(f);
}

// ----------------------------------------------------------------------------------------------------
// Synthetic periodic cube [0,L]^3, ne^3 elements.  amp != 0 curves the interior with
// x_d += amp*sin(2 pi x/L) sin(2 pi y/L) sin(2 pi z/L) (zero on the periodic planes), expressed as
// curved face patches of order bFaceOrder at Chebyshev-Lobatto points exactly as a SpecMesh file would.
// shuffle != 0 re-orients every element by a random proper rotation of its local frame (seeded), which
// exercises all eight face rotations.
inline void properRotations(std::vector<std::array<int, 6>>& out) {
    // each rotation: perm[3] (new axis a takes old axis perm[a]) and sign[3]
    int p[3] = {0, 1, 2};
    std::sort(p, p + 3);
    do {
        for (int s = 0; s < 8; ++s) {
            int sg[3] = {(s & 1) ? -1 : 1, (s & 2) ? -1 : 1, (s & 4) ? -1 : 1};
            // determinant = parity(p) * prod(sg)
            int inv = 0; for (int a = 0; a < 3; ++a) for (int b2 = a + 1; b2 < 3; ++b2) if (p[a] > p[b2]) ++inv;
            int det = ((inv % 2) ? -1 : 1) * sg[0] * sg[1] * sg[2];
            if (det == 1) out.push_back({p[0], p[1], p[2], sg[0], sg[1], sg[2]});
        }
    } while (std::next_permutation(p, p + 3));
}

inline void boxMesh(HostMesh& m, int ne, double L, double amp, int bFaceOrder, int shuffle, unsigned seed,
                    int nex = 0, int ney = 0, int nez = 0) {
    if (nex <= 0) nex = ne; if (ney <= 0) ney = ne; if (nez <= 0) nez = ne;
    const int npx = nex + 1, npy = ney + 1, npz = nez + 1;
    const double two_pi_L = 2.0 * PI_RP / L;
    auto warp = [&](double* X) {
        if (amp == 0.0) return;
        double d = amp * std::sin(two_pi_L * X[0]) * std::sin(two_pi_L * X[1]) * std::sin(two_pi_L * X[2]);
        X[0] += d; X[1] += d; X[2] += d;
    };
    m = HostMesh();
    m.bFaceOrder = (amp == 0.0) ? 1 : bFaceOrder;
    m.nodes.resize(3 * (size_t)npx * npy * npz);
    std::vector<double> straight(m.nodes.size());
    for (int c = 0; c < npz; ++c) for (int b = 0; b < npy; ++b) for (int a = 0; a < npx; ++a) {
        size_t id = a + (size_t)npx * (b + (size_t)npy * c);
        // exact end planes so that periodic matching sees identical coordinates
        double X[3] = {a == nex ? L : L * a / nex, b == ney ? L : L * b / ney, c == nez ? L : L * c / nez};
        for (int d = 0; d < 3; ++d) straight[3 * id + d] = X[d];
        if (a != 0 && a != nex && b != 0 && b != ney && c != 0 && c != nez) warp(X);
        for (int d = 0; d < 3; ++d) m.nodes[3 * id + d] = X[d];
    }
    const int nE = nex * ney * nez;
    m.elemNodes.resize(8 * (size_t)nE); m.isHex8.assign(nE, amp == 0.0 ? 1 : 0); m.patches.resize(nE); m.bname.assign(6 * (size_t)nE, "---");
    static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    std::vector<std::array<int, 6>> rots; properRotations(rots);
    std::mt19937 rng(seed);
    const int nb = m.bFaceOrder + 1;
    std::vector<double> cgl(nb);
    for (int i = 0; i < nb; ++i) cgl[i] = -std::cos(i * PI_RP / (nb - 1.0));
    for (int ez = 0; ez < nez; ++ez) for (int ey = 0; ey < ney; ++ey) for (int ex = 0; ex < nex; ++ex) {
        const int e = ex + nex * (ey + ney * ez);
        std::array<int, 6> R = {0, 1, 2, 1, 1, 1};
        if (shuffle) R = rots[rng() % rots.size()];
        // local (new) reference coords u -> physical-box reference coords s : s[R[a]] = sg[a]*u[a]
        auto localToBox = [&](const double* u, double* s) { for (int a = 0; a < 3; ++a) s[R[a]] = R[3 + a] * u[a]; };
        const int e0[3] = {ex, ey, ez};
        for (int k = 0; k < 8; ++k) {
            double u[3] = {2.0 * corner[k][0] - 1, 2.0 * corner[k][1] - 1, 2.0 * corner[k][2] - 1}, s[3];
            localToBox(u, s);
            int a = e0[0] + (s[0] > 0), b = e0[1] + (s[1] > 0), c = e0[2] + (s[2] > 0);
            m.elemNodes[8 * e + k] = a + npx * (b + npy * c);
        }
        // boundary names of the local faces
        static const char* lowName[3] = {"left", "front", "bottom"};
        static const char* highName[3] = {"right", "back", "top"};
        const int nexyz[3] = {nex, ney, nez};
        for (int f = 0; f < 6; ++f) {
            double u[3] = {0, 0, 0}, s[3];
            u[faceNormalAxis[f]] = faceNormalEnd[f] ? 1.0 : -1.0;
            localToBox(u, s);
            for (int d = 0; d < 3; ++d) {
                if (s[d] < -0.5 && e0[d] == 0) m.bname[6 * e + f] = lowName[d];
                if (s[d] > 0.5 && e0[d] == nexyz[d] - 1) m.bname[6 * e + f] = highName[d];
            }
        }
        if (amp != 0.0) {
            const double h[3] = {L / nex, L / ney, L / nez};
            for (int f = 0; f < 6; ++f) {
                FacePatch& p = m.patches[e][f];
                p.nu = p.nv = nb; p.pts.resize(3 * (size_t)nb * nb);
                for (int j = 0; j < nb; ++j) for (int i = 0; i < nb; ++i) {
                    double u[3], s[3], X[3];
                    u[axisMap[f][0]] = cgl[i]; u[axisMap[f][1]] = cgl[j]; u[faceNormalAxis[f]] = faceNormalEnd[f] ? 1.0 : -1.0;
                    localToBox(u, s);
                    for (int d = 0; d < 3; ++d) X[d] = (e0[d] + 0.5 * (s[d] + 1.0)) * h[d];
                    // snap element-boundary planes to the node coordinates used by connectivity
                    for (int d = 0; d < 3; ++d) {
                        if (s[d] == -1.0) X[d] = L * e0[d] / nexyz[d];
                        if (s[d] == 1.0) X[d] = (e0[d] + 1 == nexyz[d]) ? L : L * (e0[d] + 1) / nexyz[d];
                    }
                    bool onBoundaryPlane = false;
                    for (int d = 0; d < 3; ++d) if (X[d] == 0.0 || X[d] == L) onBoundaryPlane = true;
                    if (!onBoundaryPlane) warp(X);
                    for (int c = 0; c < 3; ++c) p.pts[(j * nb + i) * 3 + c] = X[c];
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
struct Key4 { int a[4]; bool operator==(const Key4& o) const { return a[0] == o.a[0] && a[1] == o.a[1] && a[2] == o.a[2] && a[3] == o.a[3]; } };
struct Key4Hash { size_t operator()(const Key4& k) const { uint64_t h = 1469598103934665603ull; for (int i = 0; i < 4; ++i) { h ^= (uint64_t)(uint32_t)k.a[i]; h *= 1099511628211ull; } return (size_t)h; } };

inline bool compareTwoNodes(const double* x1, const double* x2, int& coord) {
    int counter = 0;
    if (coord == 0) {
        for (int i = 1; i <= 3; ++i) { if (almostEqual(x1[i - 1], x2[i - 1])) ++counter; else coord = i; }
    } else {
        for (int i = 1; i <= 3; ++i) if (i != coord && almostEqual(x1[i - 1], x2[i - 1])) ++counter;
    }
    return counter >= 2;
}

// Builds faces, merges periodic pairs, fills element->face tables.  Returns false with err on failure.
inline bool buildConnectivity(HostMesh& m, const std::vector<BCSpec>& bcs, std::string& err) {
    m.bcs = bcs;
    const int nE = m.nElem();
    std::unordered_map<Key4, int, Key4Hash> table;
    table.reserve((size_t)nE * 4);
    std::vector<int> fNodes, fElem, fSide, fRot, fType, fZone;
    std::vector<std::string> fName;
    for (int e = 0; e < nE; ++e) for (int fn = 0; fn < 6; ++fn) {
        int ids[4]; for (int j = 0; j < 4; ++j) ids[j] = m.elemNodes[8 * e + localFaceNode[fn][j]];
        Key4 key; std::copy(ids, ids + 4, key.a); std::sort(key.a, key.a + 4);
        auto it = table.find(key);
        if (it != table.end()) {
            int f = it->second;
            if (fElem[2 * f + 1] >= 0) { err = "face shared by more than two elements"; return false; }
            fElem[2 * f + 1] = e; fSide[2 * f + 1] = fn; fType[f] = HMESH_INTERIOR;
            fRot[f] = faceRotation(&fNodes[4 * f], ids);
        } else {
            int f = (int)fRot.size();
            table.emplace(key, f);
            fNodes.insert(fNodes.end(), ids, ids + 4);
            fElem.push_back(e); fElem.push_back(-1); fSide.push_back(fn); fSide.push_back(-1);
            fRot.push_back(0); fType.push_back(HMESH_UNDEFINED); fZone.push_back(-1);
            fName.push_back(m.bname[6 * e + fn]);
        }
    }
    int nF = (int)fRot.size();
    // zones
    auto zoneOf = [&](const std::string& nm) { for (size_t z = 0; z < bcs.size(); ++z) if (bcs[z].name == nm) return (int)z; return -1; };
    std::vector<std::vector<int>> zoneFaces(bcs.size());
    for (int f = 0; f < nF; ++f) if (fType[f] == HMESH_UNDEFINED) {
        if (fName[f] == "---") { err = "unconnected face without a boundary name"; return false; }
        int z = zoneOf(fName[f]);
        if (z < 0) { err = "boundary \"" + fName[f] + "\" not defined in the boundary-condition table"; return false; }
        fZone[f] = z; zoneFaces[z].push_back(f);
    }
    // periodic pairing
    std::vector<char> zoneDeleted(bcs.size(), 0);
    for (size_t zp = 0; zp < bcs.size(); ++zp) {
        if (bcs[zp].type != "periodic" || zoneDeleted[zp]) continue;
        int zm = zoneOf(bcs[zp].coupled);
        if (zm < 0) { err = "coupled boundary \"" + bcs[zp].coupled + "\" for boundary \"" + bcs[zp].name + "\" not found."; return false; }
        zoneDeleted[zm] = 1;
        int coord = 0;
        std::vector<char> taken(zoneFaces[zm].size(), 0);
        for (int i : zoneFaces[zp]) {
            if (fType[i] != HMESH_UNDEFINED) continue;
            bool paired = false;
            for (size_t jj = 0; jj < zoneFaces[zm].size() && !paired; ++jj) {
                if (taken[jj]) continue;
                int j = zoneFaces[zm][jj];
                bool mm[4] = {false, false, false, false}, sm[4];
                int usedCoord = coord;
                auto tryCoord = [&](int lc) {
                    for (int k = 0; k < 4; ++k) { mm[k] = false; sm[k] = false; }
                    for (int k = 0; k < 4; ++k) {
                        const double* x1 = &m.nodes[3 * fNodes[4 * i + k]];
                        for (int l = 0; l < 4; ++l) if (!sm[l]) {
                            const double* x2 = &m.nodes[3 * fNodes[4 * j + l]];
                            int c = lc; if (compareTwoNodes(x1, x2, c)) { mm[k] = true; sm[l] = true; break; }
                        }
                        if (!mm[k]) return false;
                    }
                    return true;
                };
                bool ok = false;
                if (coord == 0) { for (int lc = 1; lc <= 3 && !ok; ++lc) { ok = tryCoord(lc); if (ok) usedCoord = lc; } }
                else ok = tryCoord(coord);
                if (!ok) continue;
                if (coord == 0) coord = usedCoord;
                fName[i] = "---"; fElem[2 * i + 1] = fElem[2 * j]; fSide[2 * i + 1] = fSide[2 * j]; fType[i] = HMESH_INTERIOR; fZone[i] = -1;
                m.bname[6 * fElem[2 * i] + fSide[2 * i]] = "---"; m.bname[6 * fElem[2 * i + 1] + fSide[2 * i + 1]] = "---";
                int slave[4] = {-1, -1, -1, -1};
                for (int k = 0; k < 4; ++k) for (int l = 0; l < 4; ++l) {
                    int c = coord;
                    if (compareTwoNodes(&m.nodes[3 * fNodes[4 * i + k]], &m.nodes[3 * fNodes[4 * j + l]], c)) slave[l] = fNodes[4 * i + k];
                }
                fRot[i] = faceRotation(&fNodes[4 * i], slave);
                if (fRot[i] < 0) { err = "could not determine periodic face rotation"; return false; }
                taken[jj] = 1; paired = true;
            }
            if (!paired) { err = "periodic face was not able to find a partner (zone " + bcs[zp].name + ")"; return false; }
        }
    }
    // delete periodic- faces, renumber
    m.faceNodes.clear(); m.faceElem.clear(); m.faceElemSide.clear(); m.faceRot.clear(); m.faceType.clear(); m.faceZone.clear();
    for (int f = 0; f < nF; ++f) {
        bool keep = fType[f] != HMESH_UNDEFINED || !zoneDeleted[fZone[f]];
        if (!keep) continue;
        if (fType[f] == HMESH_UNDEFINED) {
            if (bcs[fZone[f]].type == "periodic") { err = "unpaired periodic face left in zone " + bcs[fZone[f]].name; return false; }
            fType[f] = HMESH_BOUNDARY;
        }
        m.faceNodes.insert(m.faceNodes.end(), &fNodes[4 * f], &fNodes[4 * f] + 4);
        m.faceElem.push_back(fElem[2 * f]); m.faceElem.push_back(fElem[2 * f + 1]);
        m.faceElemSide.push_back(fSide[2 * f]); m.faceElemSide.push_back(fSide[2 * f + 1]);
        m.faceRot.push_back(fRot[f]); m.faceType.push_back(fType[f]); m.faceZone.push_back(fZone[f]);
    }
    m.nFaces = (int)m.faceRot.size();
    m.elemFace.assign(6 * (size_t)nE, -1); m.elemFaceSide.assign(6 * (size_t)nE, -1);
    for (int f = 0; f < m.nFaces; ++f) for (int s = 0; s < 2; ++s) {
        int e = m.faceElem[2 * f + s]; if (e < 0) continue;
        m.elemFace[6 * e + m.faceElemSide[2 * f + s]] = f; m.elemFaceSide[6 * e + m.faceElemSide[2 * f + s]] = s;
    }
    for (size_t q = 0; q < m.elemFace.size(); ++q) if (m.elemFace[q] < 0) { err = "element face without a mesh face"; return false; }
    return true;
}

}  // namespace h3d
