// C API of the host-side driver library (libh3dhost.so): mesh, connectivity, nodal operators, geometry,
// partitioning.  Consumed by the Python/ctypes front end and by C++ drivers; the arrays it produces are
// exactly the inputs of the device C-ABI (include/h3d_gpu.h).
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <string>

#include "geometry.hpp"
#include "geometry_p.hpp"
#include "gmsh.hpp"
#include "partition.hpp"

using namespace h3d;

namespace {
thread_local std::string g_err;
struct Host {
    HostMesh mesh;
    HostGeometry geom;
    HostGeometryP geomP;      // p-nonconforming meshes (geometry_p.hpp)
    HaloInfo halo;
    bool connected = false, hasGeom = false, mixed = false;
};
}  // namespace

extern "C" {

const char* h3dhost_last_error() { return g_err.c_str(); }

// Threads of the OpenMP loops of this library (metric terms, wall distances).  Launchers such as torchrun export
// OMP_NUM_THREADS=1 and the OpenMP runtime reads it once, when the first library that uses it is loaded -- changing the
// environment afterwards has no effect, this call has.  Returns the number of threads a parallel region gets.
int h3dhost_set_num_threads(int nt) {
#ifdef _OPENMP
    if (nt > 0) omp_set_num_threads(nt);
    int got = 1;
#pragma omp parallel
    {
#pragma omp single
        got = omp_get_num_threads();
    }
    return got;
#else
    (void)nt; return 1;
#endif
}

void* h3dhost_mesh_box(int nex, int ney, int nez, double L, double amp, int bFaceOrder, int shuffle, unsigned seed) {
    Host* h = new Host();
    boxMesh(h->mesh, nex, L, amp, bFaceOrder, shuffle, seed, nex, ney, nez);
    return h;
}

void* h3dhost_mesh_read(const char* path) {
    Host* h = new Host();
    // as the reference's mesh-file dispatch (ReadMeshFile.f90:37-77): the extension selects the reader
    const std::string p(path);
    const bool gmsh = p.size() > 4 && p.compare(p.size() - 4, 4, ".msh") == 0;
    if (!(gmsh ? readGmsh(p, h->mesh, g_err) : readSpecMesh(p, h->mesh, g_err))) { delete h; return nullptr; }
    return h;
}

int h3dhost_mesh_write(void* hp, const char* path) { writeSpecMesh(path, ((Host*)hp)->mesh); return 0; }

void h3dhost_mesh_free(void* hp) { delete (Host*)hp; }

// bc table: names/types/coupled are nbc C strings each; params is 16 doubles per bc
int h3dhost_mesh_connect(void* hp, int nbc, const char** names, const char** types, const char** coupled, const double* params) {
    Host* h = (Host*)hp;
    std::vector<BCSpec> bcs(nbc);
    for (int i = 0; i < nbc; ++i) {
        bcs[i].name = toLower(names[i]); bcs[i].type = toLower(types[i]); bcs[i].coupled = coupled[i] ? toLower(coupled[i]) : "";
        if (params) std::memcpy(bcs[i].params, params + 16 * i, 16 * sizeof(double));
    }
    if (!buildConnectivity(h->mesh, bcs, g_err)) return 1;
    h->connected = true;
    return 0;
}

int h3dhost_mesh_geometry(void* hp, int N, int nodeType) {
    Host* h = (Host*)hp;
    if (!h->connected) { g_err = "mesh connectivity has not been built"; return 1; }
    if (N < 1 || N > 15) { g_err = "polynomial order out of range"; return 1; }
    // nodeType + 16: the metric terms are interpolated to the nodes with the reference's full triple sum (n^6 per element, for the
    // regression pins) instead of the sum-factorised form
    const bool referenceOrder = (nodeType & 16) != 0; nodeType &= 15;
    try { buildGeometry(h->mesh, N, nodeType, h->geom, referenceOrder); } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
    h->hasGeom = true;
    return 0;
}

// Geometry of a p-nonconforming mesh: Nxyz[nElem][3] are the elements' polynomial orders (the reference's "polynomial order
// file", ReadOrderFile, libs/io/ReadInputFile.f90:132-153).  Afterwards h3dhost_get_array serves the packed arrays of
// geometry_p.hpp under the usual names, plus "elemOrder" [e][3] and "faceOrder" [f][6] = Nf, NfLeft, NfRight.
int h3dhost_mesh_geometry_p(void* hp, int nodeType, const int* Nxyz) {
    Host* h = (Host*)hp;
    if (!h->connected) { g_err = "mesh connectivity has not been built"; return 1; }
    try { if (!buildGeometryP(h->mesh, Nxyz, nodeType, h->geomP, g_err)) return 1; } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
    h->hasGeom = true; h->mixed = true;
    return 0;
}

// Tset(Norigin, Ndest) % T (libs/spectral/InterpolationMatrices.f90:42-107), row-major [(Ndest+1)][(Norigin+1)]
int h3dhost_interpolation_matrix(int Norigin, int Ndest, int nodeType, double* T) {
    try {
        NodalStorage a, b; a.construct(nodeType, Norigin); b.construct(nodeType, Ndest);
        std::vector<double> M; interpolationMatrix(a, b, M);
        std::memcpy(T, M.data(), M.size() * sizeof(double));
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
    return 0;
}

int h3dhost_wall_distance(void* hp) {
    Host* h = (Host*)hp;
    if (h->mixed) {
        if (!h->hasGeom) { g_err = "geometry has not been built"; return 1; }
        const std::vector<double> Xw = wallCoordinatesP(h->mesh, h->geomP);
        if (Xw.empty()) { g_err = "no wall points: the wall model needs at least one no-slip wall face in the whole mesh"; return 1; }
        computeWallDistancesP(h->geomP, Xw);
        return 0;
    }
    if (!h->hasGeom) { g_err = "geometry has not been built"; return 1; }
    computeWallDistances(h->mesh, h->geom);
    return 0;
}

// The two halves of the reference's parallel wall-distance computation (HexMesh.f90:5594-5780): every rank contributes the
// nodes of its no-slip wall faces (wall_points: count = number of points; pts may be NULL to ask for the count), the driver
// gathers them across ranks, and every rank measures its nodes against the whole set (wall_distance_from).
int h3dhost_wall_points(void* hp, double* pts, long long* count) {
    Host* h = (Host*)hp;
    if (h->mixed) { g_err = "wall distances are not available on p-nonconforming meshes"; return 1; }
    if (!h->hasGeom) { g_err = "geometry has not been built"; return 1; }
    const std::vector<double> Xw = wallCoordinates(h->mesh, h->geom);
    *count = (long long)(Xw.size() / 3);
    if (pts && !Xw.empty()) std::memcpy(pts, Xw.data(), Xw.size() * sizeof(double));
    return 0;
}
int h3dhost_wall_distance_from(void* hp, const double* pts, long long count) {
    Host* h = (Host*)hp;
    if (h->mixed) { g_err = "wall distances are not available on p-nonconforming meshes"; return 1; }
    if (!h->hasGeom) { g_err = "geometry has not been built"; return 1; }
    if (count <= 0) { g_err = "no wall points: the wall model needs at least one no-slip wall face in the whole mesh"; return 1; }
    computeWallDistances(h->mesh, h->geom, std::vector<double>(pts, pts + 3 * count));
    return 0;
}

int h3dhost_mesh_sizes(void* hp, int* nElem, int* nFaces, int* nNodes, int* N) {
    Host* h = (Host*)hp;
    *nElem = h->mesh.nElem(); *nFaces = h->mesh.nFaces; *nNodes = h->mesh.nNodes(); *N = h->hasGeom ? h->geom.N : -1;
    return 0;
}

// name -> pointer/count/type (0 = double, 1 = int32).  Pointers stay valid until the mesh is freed/rebuilt.
int h3dhost_get_array(void* hp, const char* name, void** ptr, long long* count, int* isInt) {
    Host* h = (Host*)hp;
    std::string s(name);
#define DARR(nm, vec) if (s == nm) { *ptr = (void*)(vec).data(); *count = (long long)(vec).size(); *isInt = 0; return 0; }
#define IARR(nm, vec) if (s == nm) { *ptr = (void*)(vec).data(); *count = (long long)(vec).size(); *isInt = 1; return 0; }
    DARR("nodes", h->mesh.nodes) IARR("elemNodes", h->mesh.elemNodes)
    IARR("faceNodes", h->mesh.faceNodes) IARR("faceElem", h->mesh.faceElem) IARR("faceElemSide", h->mesh.faceElemSide)
    IARR("faceRot", h->mesh.faceRot) IARR("faceType", h->mesh.faceType) IARR("faceZone", h->mesh.faceZone)
    IARR("elemFace", h->mesh.elemFace) IARR("elemFaceSide", h->mesh.elemFaceSide)
    if (h->mixed) {
        HostGeometryP& G = h->geomP;
        IARR("elemOrder", G.elemOrder) IARR("faceOrder", G.faceOrder)
        DARR("x", G.x) DARR("jGradXi", G.jGradXi) DARR("jGradEta", G.jGradEta) DARR("jGradZeta", G.jGradZeta)
        DARR("jacobian", G.jac) DARR("invJacobian", G.invJac) DARR("volume", G.volume)
        DARR("faceX", G.fx) DARR("faceNormal", G.fnormal) DARR("faceT1", G.ft1) DARR("faceT2", G.ft2)
        DARR("faceJacobian", G.fjac) DARR("faceSurface", G.fsurface) DARR("faceH", G.fh) DARR("dWall", G.dWall) DARR("faceDWall", G.fdWall)
    }
    DARR("x", h->geom.x) DARR("jGradXi", h->geom.jGradXi) DARR("jGradEta", h->geom.jGradEta) DARR("jGradZeta", h->geom.jGradZeta)
    DARR("jacobian", h->geom.jac) DARR("invJacobian", h->geom.invJac) DARR("volume", h->geom.volume)
    DARR("faceX", h->geom.fx) DARR("faceNormal", h->geom.fnormal) DARR("faceT1", h->geom.ft1) DARR("faceT2", h->geom.ft2)
    DARR("faceJacobian", h->geom.fjac) DARR("faceSurface", h->geom.fsurface) DARR("faceH", h->geom.fh) DARR("dWall", h->geom.dWall) DARR("faceDWall", h->geom.fdWall)
    IARR("haloRank", h->halo.rank) IARR("haloCount", h->halo.count) IARR("haloFace", h->halo.face) IARR("haloSide", h->halo.side)
    IARR("globalElem", h->halo.globalElem) IARR("globalFace", h->halo.globalFace)
#undef DARR
#undef IARR
    g_err = "unknown array: " + s;
    return 1;
}

// 1-D operators: x,w (n); D,hatD,sharpD (n*n row-major M(i,l)); v,b (2*n: side-major)
int h3dhost_nodal(int N, int nodeType, double* x, double* w, double* D, double* hatD, double* sharpD, double* v, double* b) {
    try {
        NodalStorage sp; sp.construct(nodeType, N);
        const int n = N + 1;
        std::memcpy(x, sp.x.data(), n * sizeof(double)); std::memcpy(w, sp.w.data(), n * sizeof(double));
        std::memcpy(D, sp.D.data(), n * n * sizeof(double)); std::memcpy(hatD, sp.hatD.data(), n * n * sizeof(double));
        std::memcpy(sharpD, sp.sharpD.data(), n * n * sizeof(double));
        std::memcpy(v, sp.v.data(), 2 * n * sizeof(double)); std::memcpy(b, sp.b.data(), 2 * n * sizeof(double));
    } catch (const std::exception& ex) { g_err = ex.what(); return 1; }
    return 0;
}

// Element partition: method 0 = METIS_PartMeshDual (ncommon = 4, as METISPartitioning.f90:151), 1 = contiguous
// blocks of the element list.  part[nElem] receives 0-based ranks.
int h3dhost_partition(void* hp, int nparts, int method, int* part) {
    Host* h = (Host*)hp;
    return partitionElements(h->mesh, nparts, method, part, g_err) ? 0 : 1;
}

int h3dhost_partition_weighted(void* hp, int nparts, int method, const int* vwgt, int* part) {
    Host* h = (Host*)hp;
    return partitionElements(h->mesh, nparts, method, part, g_err, vwgt) ? 0 : 1;
}

// Extracts the local mesh of `rank`: returns a new handle whose faces on partition cuts are HMESH_MPI
// (single local side, rotation kept), with geometry rebuilt when the parent has it.  Halo tables are
// exposed through h3dhost_get_array on the child: "haloRank","haloCount","haloFace","haloSide","globalElem".
void* h3dhost_extract_partition(void* hp, const int* part, int rank) {
    Host* h = (Host*)hp;
    if (!h->connected) { g_err = "mesh connectivity has not been built"; return nullptr; }
    Host* c = new Host();
    extractPartition(h->mesh, part, rank, c->mesh, c->halo);
    c->connected = true;
    return c;
}

// Copies the parent's element/face geometry into a partition extracted from it (instead of rebuilding it from the local
// elements).  MPI faces then carry the GLOBAL face geometry on both ranks, which makes results independent of the
// partition bit for bit (the reference rebuilds MPI-face geometry from the local element, HexMesh.f90:3000-3030).
int h3dhost_inherit_geometry(void* childp, void* parentp) {
    Host* c = (Host*)childp; Host* p = (Host*)parentp;
    if (!p->hasGeom) { g_err = "parent mesh has no geometry"; return 1; }
    if (p->mixed) {   // p-nonconforming: orders and packed arrays of the local elements / faces, cut faces keep the global face's orders
        const HostGeometryP& G = p->geomP; HostGeometryP& g = c->geomP;
        g = HostGeometryP(); g.nodeType = G.nodeType; g.meshIs2D = G.meshIs2D; g.anisotropic = G.anisotropic; g.sp = G.sp;
        const size_t nE = c->halo.globalElem.size(), nF = c->halo.globalFace.size();
        g.eOff.assign(nE + 1, 0); g.fOff.assign(nF + 1, 0);
        for (size_t l = 0; l < nE; ++l) {
            const int e = c->halo.globalElem[l];
            g.elemOrder.insert(g.elemOrder.end(), &G.elemOrder[3 * (size_t)e], &G.elemOrder[3 * (size_t)e] + 3);
            g.eOff[l + 1] = g.eOff[l] + (G.eOff[e + 1] - G.eOff[e]);
            g.volume.push_back(G.volume[e]);
        }
        for (size_t l = 0; l < nF; ++l) {
            const int f = c->halo.globalFace[l];
            g.faceOrder.insert(g.faceOrder.end(), &G.faceOrder[6 * (size_t)f], &G.faceOrder[6 * (size_t)f] + 6);
            g.fOff[l + 1] = g.fOff[l] + (G.fOff[f + 1] - G.fOff[f]);
            g.fsurface.push_back(G.fsurface[f]); g.fh.push_back(G.fh[f]);
        }
        auto gather = [&](const std::vector<double>& src, std::vector<double>& dst, const std::vector<long long>& offG, const std::vector<long long>& offL,
                          const std::vector<int>& ids, size_t w) {
            dst.resize((size_t)offL.back() * w);
            for (size_t l = 0; l < ids.size(); ++l)
                std::memcpy(&dst[(size_t)offL[l] * w], &src[(size_t)offG[ids[l]] * w], (size_t)(offL[l + 1] - offL[l]) * w * sizeof(double));
        };
        gather(G.x, g.x, G.eOff, g.eOff, c->halo.globalElem, 3); gather(G.jGradXi, g.jGradXi, G.eOff, g.eOff, c->halo.globalElem, 3);
        gather(G.jGradEta, g.jGradEta, G.eOff, g.eOff, c->halo.globalElem, 3); gather(G.jGradZeta, g.jGradZeta, G.eOff, g.eOff, c->halo.globalElem, 3);
        gather(G.jac, g.jac, G.eOff, g.eOff, c->halo.globalElem, 1); gather(G.invJac, g.invJac, G.eOff, g.eOff, c->halo.globalElem, 1);
        gather(G.fx, g.fx, G.fOff, g.fOff, c->halo.globalFace, 3); gather(G.fnormal, g.fnormal, G.fOff, g.fOff, c->halo.globalFace, 3);
        gather(G.ft1, g.ft1, G.fOff, g.fOff, c->halo.globalFace, 3); gather(G.ft2, g.ft2, G.fOff, g.fOff, c->halo.globalFace, 3);
        gather(G.fjac, g.fjac, G.fOff, g.fOff, c->halo.globalFace, 1);
        if (!G.dWall.empty()) { gather(G.dWall, g.dWall, G.eOff, g.eOff, c->halo.globalElem, 1); gather(G.fdWall, g.fdWall, G.fOff, g.fOff, c->halo.globalFace, 1); }
        c->hasGeom = true; c->mixed = true;
        return 0;
    }
    const HostGeometry& G = p->geom; HostGeometry& g = c->geom;
    g.N = G.N; g.n = G.n; g.nodeType = G.nodeType; g.sp = G.sp;
    const size_t n3 = (size_t)G.n * G.n * G.n, n2 = (size_t)G.n * G.n;
    const size_t nE = c->halo.globalElem.size(), nF = c->halo.globalFace.size();
    auto gatherE = [&](const std::vector<double>& src, std::vector<double>& dst, size_t w) {
        dst.resize(nE * w);
        for (size_t l = 0; l < nE; ++l) std::memcpy(&dst[l * w], &src[(size_t)c->halo.globalElem[l] * w], w * sizeof(double));
    };
    auto gatherF = [&](const std::vector<double>& src, std::vector<double>& dst, size_t w) {
        dst.resize(nF * w);
        for (size_t l = 0; l < nF; ++l) std::memcpy(&dst[l * w], &src[(size_t)c->halo.globalFace[l] * w], w * sizeof(double));
    };
    gatherE(G.x, g.x, 3 * n3); gatherE(G.jGradXi, g.jGradXi, 3 * n3); gatherE(G.jGradEta, g.jGradEta, 3 * n3); gatherE(G.jGradZeta, g.jGradZeta, 3 * n3);
    gatherE(G.jac, g.jac, n3); gatherE(G.invJac, g.invJac, n3); gatherE(G.volume, g.volume, 1);
    gatherF(G.fx, g.fx, 3 * n2); gatherF(G.fnormal, g.fnormal, 3 * n2); gatherF(G.ft1, g.ft1, 3 * n2); gatherF(G.ft2, g.ft2, 3 * n2);
    gatherF(G.fjac, g.fjac, n2); gatherF(G.fsurface, g.fsurface, 1); gatherF(G.fh, g.fh, 1);
    if (!G.dWall.empty()) { gatherE(G.dWall, g.dWall, n3); gatherF(G.fdWall, g.fdWall, n2); }   // wall distances of the global mesh
    c->hasGeom = true;
    return 0;
}

}  // extern "C"
