// Element-wise domain decomposition of the host mesh and extraction of per-rank local meshes whose cut
// faces become HMESH_MPI faces with one local side.
//
// Reference behaviour followed (paths relative to /root/reference/Solver/src/libs):
//   mesh/METISPartitioning.f90:125-151,210-251   METIS_PartMeshDual(ne, nn, eptr, eind, ncommon = 4, nparts)
//   mesh/MeshPartitioning.f90:387                 space-filling / contiguous split (method 1 here: contiguous blocks)
//   mesh/HexMesh.f90:2571-2668                    UpdateFacesWithPartition: MPI faces keep rotation and the
//                                                 side (1/2) the local element had on the global face
//   mesh/HexMesh.f90:2663-2664                    per-neighbour face lists define the exchange order
#pragma once
#include <cstdint>
#include <numeric>

#include "mesh.hpp"

extern "C" {
// METIS 5 (libmetis_static.a shipped with the CUDA toolkit): that build uses idx_t = int64 (probed), real_t = float
int METIS_SetDefaultOptions(int64_t* options);
int METIS_PartMeshDual(int64_t* ne, int64_t* nn, int64_t* eptr, int64_t* eind, int64_t* vwgt, int64_t* vsize, int64_t* ncommon, int64_t* nparts,
                       float* tpwgts, int64_t* options, int64_t* objval, int64_t* epart, int64_t* npart);
}

namespace h3d {

struct HaloInfo {
    std::vector<int> rank, count, face, side;   // face/side concatenated per neighbour in `rank` order
    std::vector<int> globalElem, globalFace;
};

// vwgt (may be null): one weight per element -- the reference passes the degrees of freedom of every element when the polynomial
// orders are not uniform (METISPartitioning.f90:125-151)
inline bool partitionElements(const HostMesh& m, int nparts, int method, int* part, std::string& err, const int* vwgt = nullptr) {
    const int nE = m.nElem();
    if (nparts < 1) { err = "nparts must be >= 1"; return false; }
    if (nparts == 1) { std::fill(part, part + nE, 0); return true; }
    if (method == 1) {
        if (!vwgt) { for (int e = 0; e < nE; ++e) part[e] = (int)(((long long)e * nparts) / nE); return true; }
        long long total = 0, acc = 0;                        // contiguous blocks of (nearly) equal weight
        for (int e = 0; e < nE; ++e) total += vwgt[e];
        for (int e = 0; e < nE; ++e) { part[e] = (int)std::min<long long>(nparts - 1, (acc * nparts) / total); acc += vwgt[e]; }
        return true;
    }
#ifdef H3D_HAS_METIS
    int64_t ne = nE, nn = m.nNodes(), ncommon = 4, np = nparts, objval = 0;
    std::vector<int64_t> eptr(nE + 1), eind(m.elemNodes.begin(), m.elemNodes.end()), npart(nn), epart(nE), options(40);
    for (int e = 0; e <= nE; ++e) eptr[e] = 8 * (int64_t)e;
    METIS_SetDefaultOptions(options.data());
    std::vector<int64_t> w; if (vwgt) w.assign(vwgt, vwgt + nE);
    int rc = METIS_PartMeshDual(&ne, &nn, eptr.data(), eind.data(), vwgt ? w.data() : nullptr, nullptr, &ncommon, &np, nullptr, options.data(), &objval, epart.data(), npart.data());
    if (rc != 1) { err = "METIS_PartMeshDual failed"; return false; }
    for (int e = 0; e < nE; ++e) part[e] = (int)epart[e];
    return true;
#else
    err = "library built without METIS";
    return false;
#endif
}

// Builds the local mesh of `rank`.  Periodic faces that METIS cut are treated like any other cut face.
inline void extractPartition(const HostMesh& g, const int* part, int rank, HostMesh& loc, HaloInfo& halo) {
    const int nE = g.nElem();
    std::vector<int> g2l(nE, -1);
    loc = HostMesh();
    loc.bFaceOrder = g.bFaceOrder; loc.nodes = g.nodes; loc.bcs = g.bcs;
    for (int e = 0; e < nE; ++e) if (part[e] == rank) {
        g2l[e] = (int)halo.globalElem.size(); halo.globalElem.push_back(e);
    }
    const int nL = (int)halo.globalElem.size();
    loc.elemNodes.resize(8 * (size_t)nL); loc.isHex8.resize(nL); loc.patches.resize(nL); loc.bname.resize(6 * (size_t)nL);
    for (int l = 0; l < nL; ++l) {
        int e = halo.globalElem[l];
        std::copy(&g.elemNodes[8 * e], &g.elemNodes[8 * e] + 8, &loc.elemNodes[8 * l]);
        loc.isHex8[l] = g.isHex8[e]; if (!g.isHex8[e]) loc.patches[l] = g.patches[e];
        for (int f = 0; f < 6; ++f) loc.bname[6 * l + f] = g.bname[6 * e + f];
    }
    std::map<int, std::vector<std::pair<int, int>>> byRank;   // neighbour -> (global face, local face)
    for (int f = 0; f < g.nFaces; ++f) {
        int eL = g.faceElem[2 * f], eR = g.faceElem[2 * f + 1];
        bool hasL = eL >= 0 && part[eL] == rank, hasR = eR >= 0 && part[eR] == rank;
        if (!hasL && !hasR) continue;
        int lf = (int)loc.faceRot.size();
        loc.faceNodes.insert(loc.faceNodes.end(), &g.faceNodes[4 * f], &g.faceNodes[4 * f] + 4);
        loc.faceElem.push_back(hasL ? g2l[eL] : -1); loc.faceElem.push_back(hasR ? g2l[eR] : -1);
        loc.faceElemSide.push_back(g.faceElemSide[2 * f]); loc.faceElemSide.push_back(g.faceElemSide[2 * f + 1]);
        loc.faceRot.push_back(g.faceRot[f]); loc.faceZone.push_back(g.faceZone[f]);
        int type = g.faceType[f];
        if (type == HMESH_INTERIOR && !(hasL && hasR)) {
            type = HMESH_MPI;
            byRank[part[hasL ? eR : eL]].push_back({f, lf});
        }
        loc.faceType.push_back(type);
        halo.globalFace.push_back(f);
    }
    loc.nFaces = (int)loc.faceRot.size();
    loc.elemFace.assign(6 * (size_t)nL, -1); loc.elemFaceSide.assign(6 * (size_t)nL, -1);
    for (int f = 0; f < loc.nFaces; ++f) for (int s = 0; s < 2; ++s) {
        int e = loc.faceElem[2 * f + s]; if (e < 0) continue;
        loc.elemFace[6 * e + loc.faceElemSide[2 * f + s]] = f; loc.elemFaceSide[6 * e + loc.faceElemSide[2 * f + s]] = s;
    }
    for (auto& kv : byRank) {
        std::sort(kv.second.begin(), kv.second.end());   // global face order: identical on both ranks
        halo.rank.push_back(kv.first); halo.count.push_back((int)kv.second.size());
        for (auto& p : kv.second) { halo.face.push_back(p.second); halo.side.push_back(loc.faceElem[2 * p.second] >= 0 ? 0 : 1); }
    }
}

}  // namespace h3d
