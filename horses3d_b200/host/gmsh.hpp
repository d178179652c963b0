// GMSH 4.1 (ASCII) reader for hexahedral meshes of order 1 and 2, producing the same raw HostMesh as the SpecMesh reader.
// Host-side test stand-in for the reference's ConstructMesh_FromGMSHFile_v4_ (libs/mesh/Read_GMSH.f90:203-842); needed for
// BASELINE configs[4] (tutorials/Cylinder/MESH/cyl_circ.msh, 27-node hexahedra).
//
// Reference behaviour followed:
//   Read_GMSH.f90:291-322   $PhysicalNames: only 2-D physical groups are boundary conditions
//   Read_GMSH.f90:330-400   $Entities: physical tags and bounding entities of curves and surfaces (signs dropped)
//   Read_GMSH.f90:402-455   $Nodes / $Elements blocks (node tags must be 1..N)
//   Read_GMSH.f90:457-483   the mesh order is that of its (single) hexahedron type; vertices re-ordered by ReorderElement (:4269-4290):
//                           HORSES corner 1 = GMSH vertex 4, corners 2-4 = GMSH 1-3, corner 5 = GMSH 8, corners 6-8 = GMSH 5-7
//   Read_GMSH.f90:485-576   node set of a boundary condition: nodes classified on its surfaces, their curves, their points
//   Read_GMSH.f90:578-606   an element face belongs to the boundary condition when its four corners are in that set
//                           (later boundary conditions overwrite earlier ones)
//   Read_GMSH.f90:620-690   curved meshes: all six faces of every element get a patch on EQUISPACED knots -1 + 2 i / order; the
//                           node of patch position (i,j) is given by GetOrderedFaceNodeTags (:4084-4265)
// For order 2 the equispaced knots (-1, 0, 1) are the Chebyshev-Lobatto knots the FacePatch of this stand-in assumes, so the
// patches carry over unchanged; orders above 2 are refused.
#pragma once
#include <set>

#include "mesh.hpp"

namespace h3d {

// GetOrderedFaceNodeTags at order 2: 1-based (re-ordered) element node of patch position i + 3 j, per local face
static const int gmshFaceNodes27[6][9] = {
    {1, 10, 2, 16, 23, 11, 5, 18, 6},  {4, 12, 3, 15, 24, 13, 8, 19, 7}, {1, 10, 2, 14, 21, 9, 4, 12, 3},
    {2, 9, 3, 11, 22, 13, 6, 17, 7},   {5, 18, 6, 20, 26, 17, 8, 19, 7}, {1, 14, 4, 16, 25, 15, 5, 20, 8}};
// corners of the local faces in the order the boundary test uses (west, east, south, front, north, back = faces 1..6)
static const int gmshFaceCorners[6][4] = {{1, 2, 5, 6}, {3, 4, 7, 8}, {1, 2, 3, 4}, {2, 3, 6, 7}, {5, 6, 7, 8}, {1, 4, 5, 8}};

inline bool readGmsh(const std::string& path, HostMesh& m, std::string& err) {
    std::ifstream in(path);
    if (!in) { err = "Error opening file: " + path; return false; }
    std::string line;
    auto expect = [&](const char* what) { return std::getline(in, line) && line.rfind(what, 0) == 0; };
    auto numbers = [&](std::vector<double>& out) {
        out.clear();
        if (!std::getline(in, line)) return false;
        std::istringstream ss(line); double v;
        while (ss >> v) out.push_back(v);
        return true;
    };
    std::vector<double> v;
    if (!expect("$MeshFormat") || !numbers(v) || v.size() < 3) { err = "READ_GMSH :: Wrong input file."; return false; }
    if ((int)v[0] != 4 || (int)v[1] != 0) { err = "READ_GMSH :: only ASCII files of format 4.x are supported by this reader"; return false; }
    if (!expect("$EndMeshFormat")) { err = "READ_GMSH :: Wrong input file."; return false; }
    // ---- boundary conditions = 2-D physical names
    if (!expect("$PhysicalNames")) { err = "READ_GMSH :: Wrong input file - no boundary conditions defined."; return false; }
    struct BC { int tag; std::string name; std::set<int> surf, curve, point, nodes; };
    std::vector<BC> bcs;
    if (!numbers(v)) { err = "READ_GMSH :: bad $PhysicalNames"; return false; }
    for (int i = 0, n = (int)v[0]; i < n; ++i) {
        if (!std::getline(in, line)) { err = "READ_GMSH :: bad $PhysicalNames"; return false; }
        std::istringstream ss(line); int dim, tag; std::string name;
        ss >> dim >> tag; std::getline(ss, name);
        const size_t a = name.find('"'), b = name.rfind('"');
        if (a != std::string::npos && b > a) name = name.substr(a + 1, b - a - 1);
        if (dim == 2) bcs.push_back({tag, toLower(name), {}, {}, {}, {}});
    }
    if (!expect("$EndPhysicalNames")) { err = "READ_GMSH :: Wrong input file - not all boundary conditions detected."; return false; }
    // ---- entities
    if (!expect("$Entities") || !numbers(v) || v.size() < 4) { err = "READ_GMSH :: Wrong input file - no entities found."; return false; }
    const int nP = (int)v[0], nC = (int)v[1], nS = (int)v[2], nV = (int)v[3];
    struct Ent { int tag; std::vector<int> ptags, bps; };
    std::vector<Ent> curves(nC), surfs(nS);
    for (int i = 0; i < nP; ++i) if (!numbers(v)) { err = "READ_GMSH :: bad point entity"; return false; }
    auto readEnt = [&](Ent& e) {
        if (!numbers(v) || v.size() < 8) return false;
        e.tag = (int)v[0];
        const int np = (int)v[7];
        for (int q = 0; q < np; ++q) e.ptags.push_back((int)v[8 + q]);
        const int nb = (int)v[8 + np];
        for (int q = 0; q < nb; ++q) e.bps.push_back(std::abs((int)v[9 + np + q]));
        return true;
    };
    for (auto& c : curves) if (!readEnt(c)) { err = "READ_GMSH :: bad curve entity"; return false; }
    for (auto& s : surfs) if (!readEnt(s)) { err = "READ_GMSH :: bad surface entity"; return false; }
    for (int i = 0; i < nV; ++i) if (!numbers(v)) { err = "READ_GMSH :: bad volume entity"; return false; }
    if (!expect("$EndEntities")) { err = "READ_GMSH :: Wrong input file - not all entities detected."; return false; }
    for (auto& bc : bcs) {
        for (const auto& s : surfs) if (std::find(s.ptags.begin(), s.ptags.end(), bc.tag) != s.ptags.end()) { bc.surf.insert(s.tag); bc.curve.insert(s.bps.begin(), s.bps.end()); }
        for (const auto& c : curves) if (bc.curve.count(c.tag)) bc.point.insert(c.bps.begin(), c.bps.end());
    }
    // ---- nodes
    if (!expect("$Nodes") || !numbers(v) || v.size() < 4) { err = "READ_GMSH :: Wrong input file - no nodes found."; return false; }
    const int nBlocks = (int)v[0], nNodes = (int)v[1];
    if (nNodes != (int)v[3]) { err = "READ_gmsh :: Incoherent node numbering."; return false; }
    m.nodes.assign(3 * (size_t)nNodes, 0.0);
    for (int b = 0; b < nBlocks; ++b) {
        if (!numbers(v) || v.size() < 4) { err = "READ_GMSH :: bad node block"; return false; }
        const int dim = (int)v[0], etag = (int)v[1], cnt = (int)v[3];
        if ((int)v[2] != 0) { err = "READ_gmsh :: Parametric nodes not supported."; return false; }
        std::vector<int> tags(cnt);
        for (int q = 0; q < cnt; ++q) { if (!numbers(v) || v.empty()) { err = "READ_GMSH :: bad node tag"; return false; } tags[q] = (int)v[0]; }
        for (int q = 0; q < cnt; ++q) {
            if (!numbers(v) || v.size() < 3 || tags[q] < 1 || tags[q] > nNodes) { err = "READ_GMSH :: bad node coordinates"; return false; }
            for (int c = 0; c < 3; ++c) m.nodes[3 * (size_t)(tags[q] - 1) + c] = v[c];
        }
        for (auto& bc : bcs) {
            const std::set<int>& where = dim == 0 ? bc.point : (dim == 1 ? bc.curve : bc.surf);
            if (dim <= 2 && where.count(etag)) bc.nodes.insert(tags.begin(), tags.end());
        }
    }
    if (!expect("$EndNodes")) { err = "READ_GMSH :: Wrong input file - not all nodes detected."; return false; }
    // ---- elements: keep the hexahedra (type 5: 8 nodes, type 12: 27 nodes); one type per mesh
    if (!expect("$Elements") || !numbers(v) || v.size() < 4) { err = "READ_GMSH :: Wrong input file - no elements found."; return false; }
    const int nEB = (int)v[0];
    std::vector<std::vector<int>> hexes; int hexType = 0;
    for (int b = 0; b < nEB; ++b) {
        if (!numbers(v) || v.size() < 4) { err = "READ_GMSH :: bad element block"; return false; }
        const int type = (int)v[2], cnt = (int)v[3];
        for (int q = 0; q < cnt; ++q) {
            if (!numbers(v)) { err = "READ_GMSH :: bad element"; return false; }
            if (type == 5 || type == 12) {
                if (hexType && hexType != type) { err = "READ_GMSH :: More than 1 type of hexahedral detected in the mesh."; return false; }
                hexType = type;
                std::vector<int> el(v.size() - 1);
                for (size_t k = 1; k < v.size(); ++k) el[k - 1] = (int)v[k];
                hexes.push_back(el);
            } else if (type == 92 || type == 93) { err = "READ_GMSH :: hexahedra of order above 2 are not supported by this reader"; return false; }
        }
    }
    if (!hexType) { err = "READ_GMSH :: No 3D elements detected in the mesh."; return false; }
    const int order = hexType == 5 ? 1 : 2, nPer = hexType == 5 ? 8 : 27;
    const int nE = (int)hexes.size();
    m.bFaceOrder = order;
    m.elemNodes.resize(8 * (size_t)nE); m.isHex8.assign(nE, order == 1); m.patches.resize(nE); m.bname.assign(6 * (size_t)nE, "---");
    for (int l = 0; l < nE; ++l) {
        std::vector<int>& el = hexes[l];
        if ((int)el.size() != nPer) { err = "READ_GMSH :: wrong number of nodes in a hexahedron"; return false; }
        const std::vector<int> g = el;   // ReorderElement: vertices only at orders 1 and 2
        el[0] = g[3]; el[1] = g[0]; el[2] = g[1]; el[3] = g[2]; el[4] = g[7]; el[5] = g[4]; el[6] = g[5]; el[7] = g[6];
        for (int k = 0; k < 8; ++k) m.elemNodes[8 * (size_t)l + k] = el[k] - 1;
        if (order == 2)
            for (int k = 0; k < 6; ++k) {
                FacePatch& p = m.patches[l][k];
                p.nu = p.nv = 3; p.pts.resize(27);
                for (int q = 0; q < 9; ++q) for (int c = 0; c < 3; ++c) p.pts[3 * q + c] = m.nodes[3 * (size_t)(el[gmshFaceNodes27[k][q] - 1] - 1) + c];
            }
        for (const auto& bc : bcs)
            for (int k = 0; k < 6; ++k) {
                int hit = 0;
                for (int q = 0; q < 4; ++q) hit += (int)bc.nodes.count(el[gmshFaceCorners[k][q] - 1]);
                if (hit == 4) m.bname[6 * (size_t)l + k] = bc.name;
            }
    }
    return true;
}

}  // namespace h3d
