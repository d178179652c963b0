// Host-side metric terms: element mapping, conservative (curl-form) metrics, face normals/tangents.
// This is driver-side pre-processing (the reference keeps it in Fortran); the device receives the arrays.
//
// Reference behaviour followed (paths relative to /root/reference/Solver/src/libs/mesh):
//   TransfiniteMaps3D.f90:186-281   Hex8 map and gradient
//   TransfiniteMaps3D.f90:300-430   general transfinite map from six face patches
//   FacePatchClass.f90:189-268      patch evaluation (bilinear shortcut for 2x2 patches)
//   MappedGeometry.f90:82-152       node coordinates, volume
//   MappedGeometry.f90:174-384      curl-form metrics on Chebyshev-Lobatto points, interpolated to the nodes
//   MappedGeometry.f90:482-756      face normal, surface Jacobian, tangents (from the LEFT element, HexMesh.f90:2990-3030)
//   TransfiniteMaps3D.f90:446-745   analytic derivative of the blend (GradGeneralHexTransfiniteMap), FacePatchClass.f90:312-417
// Deviation (round-off level only): the CGL->node interpolation of the metric terms is sum-factorised (n^4 instead of the
// reference's n^6 operations per element); `referenceOrder` restores the reference's triple sum for the regression pins.
#pragma once
#include <algorithm>
#include <cstring>

#include "mesh.hpp"

namespace h3d {

struct HostGeometry {
    int N = 0, n = 0, nodeType = GAUSS;
    NodalStorage sp;
    // element arrays in the reference's per-element order: [e][k][j][i][c]
    std::vector<double> x, jGradXi, jGradEta, jGradZeta;   // 3*n^3 per element
    std::vector<double> jac, invJac;                        // n^3 per element
    std::vector<double> volume;                             // per element
    // face arrays: [f][j][i][c]
    std::vector<double> fx, fnormal, ft1, ft2;              // 3*n^2 per face
    std::vector<double> fjac;                               // n^2 per face
    std::vector<double> fsurface;                           // per face
    std::vector<double> fh;                                 // per face: f % geom % h (HexMesh.f90:3020-3040)
    std::vector<double> dWall, fdWall;                      // n^3 per element, n^2 per face (optional)
};

inline void computeFaceMinimumDistance(const HostMesh& m, HostGeometry& g);

struct ElemMap {
    const HostMesh* m; int e; double corners[8][3];
    std::vector<double> wb[6][2];  // barycentric weights of the patch knots
    std::vector<double> knots[6][2];
    std::vector<double> Dk[6][2];  // derivative matrices on the patch knots (FacePatch % Du, Dv)
    void init(const HostMesh& mesh, int e_) {
        m = &mesh; e = e_;
        for (int k = 0; k < 8; ++k) for (int c = 0; c < 3; ++c) corners[k][c] = mesh.nodes[3 * mesh.elemNodes[8 * e + k] + c];
        if (!mesh.isHex8[e]) for (int f = 0; f < 6; ++f) {
            const FacePatch& p = mesh.patches[e][f];
            const int nn[2] = {p.nu, p.nv};
            for (int d = 0; d < 2; ++d) {
                knots[f][d].resize(nn[d]); wb[f][d].resize(nn[d]);
                for (int i = 0; i < nn[d]; ++i) knots[f][d][i] = (nn[d] == 2) ? (i ? 1.0 : -1.0) : -std::cos(i * PI_RP / (nn[d] - 1.0));
                barycentricWeights(nn[d] - 1, knots[f][d].data(), wb[f][d].data());
                Dk[f][d].resize((size_t)nn[d] * nn[d]);
                polynomialDerivativeMatrix(nn[d] - 1, knots[f][d].data(), Dk[f][d].data());
            }
        }
    }
    // ComputeFaceDerivative (FacePatchClass.f90:312-351, Compute2DPolyDeriv :364-417): g[c][0] = d x_c / du, g[c][1] = d x_c / dv
    void faceDerivative(int f, double u, double v, double g[3][2]) const {
        const FacePatch& fp = m->patches[e][f];
        if (fp.nu == 2 && fp.nv == 2) {
            for (int c = 0; c < 3; ++c) {
                const double p11 = fp.pts[0 * 3 + c], p21 = fp.pts[1 * 3 + c], p22 = fp.pts[3 * 3 + c], p12 = fp.pts[2 * 3 + c];
                g[c][0] = 0.25 * (-p11 * (1.0 - v) + p21 * (1.0 - v) + p22 * (1.0 + v) - p12 * (1.0 + v));
                g[c][1] = 0.25 * (-p11 * (1.0 - u) - p21 * (1.0 + u) + p22 * (1.0 + u) + p12 * (1.0 - u));
            }
            return;
        }
        double li[32], lj[32], dli[32], dlj[32];
        interpolatingPolynomialVector(u, fp.nu - 1, knots[f][0].data(), wb[f][0].data(), li);
        interpolatingPolynomialVector(v, fp.nv - 1, knots[f][1].data(), wb[f][1].data(), lj);
        for (int i = 0; i < fp.nu; ++i) { dli[i] = 0.0; for (int j = 0; j < fp.nu; ++j) dli[i] += Dk[f][0][j * fp.nu + i] * li[j]; }
        for (int i = 0; i < fp.nv; ++i) { dlj[i] = 0.0; for (int j = 0; j < fp.nv; ++j) dlj[i] += Dk[f][1][j * fp.nv + i] * lj[j]; }
        for (int c = 0; c < 3; ++c) g[c][0] = g[c][1] = 0.0;
        for (int j = 0; j < fp.nv; ++j) for (int i = 0; i < fp.nu; ++i)
            for (int c = 0; c < 3; ++c) {
                const double pt = fp.pts[(j * fp.nu + i) * 3 + c];
                g[c][0] += pt * dli[i] * lj[j];
                g[c][1] += pt * li[i] * dlj[j];
            }
    }
    // GradGeneralHexTransfiniteMap + ComputeGradHexTransfiniteMap (TransfiniteMaps3D.f90:446-600, 617-745): the analytic derivative of
    // the blend of the six face patches, gg[i][j] = d x_i / d xi_j; entries below 100 eps are set to zero as the reference does
    void gradGeneral(const double* u, double gg[3][3]) const {
        double face[6][3], edge[12][3], fd[6][3][2], ed[12][3], g2[3][2];
        facePoint(0, u[0], -1.0, edge[0]); facePoint(0, 1.0, u[2], edge[1]); facePoint(0, u[0], 1.0, edge[2]); facePoint(0, -1.0, u[2], edge[3]);
        facePoint(1, u[0], -1.0, edge[4]); facePoint(1, 1.0, u[2], edge[5]); facePoint(1, u[0], 1.0, edge[6]); facePoint(1, -1.0, u[2], edge[7]);
        facePoint(3, u[1], -1.0, edge[9]); facePoint(5, u[1], -1.0, edge[8]); facePoint(3, u[1], 1.0, edge[10]); facePoint(5, u[1], 1.0, edge[11]);
        facePoint(0, u[0], u[2], face[0]); facePoint(1, u[0], u[2], face[1]); facePoint(2, u[0], u[1], face[2]);
        facePoint(3, u[1], u[2], face[3]); facePoint(4, u[0], u[1], face[4]); facePoint(5, u[1], u[2], face[5]);
        auto edgeDer = [&](int k, int f, double a, double b, int which) {
            faceDerivative(f, a, b, g2);
            for (int c = 0; c < 3; ++c) ed[k][c] = 2.0 * g2[c][which];
        };
        edgeDer(0, 0, u[0], -1.0, 0); edgeDer(1, 0, 1.0, u[2], 1); edgeDer(2, 0, u[0], 1.0, 0); edgeDer(3, 0, -1.0, u[2], 1);
        edgeDer(4, 1, u[0], -1.0, 0); edgeDer(5, 1, 1.0, u[2], 1); edgeDer(6, 1, u[0], 1.0, 0); edgeDer(7, 1, -1.0, u[2], 1);
        edgeDer(8, 5, u[1], -1.0, 0); edgeDer(9, 3, u[1], -1.0, 0); edgeDer(10, 3, u[1], 1.0, 0); edgeDer(11, 5, u[1], 1.0, 0);
        const double fu[6][2] = {{u[0], u[2]}, {u[0], u[2]}, {u[0], u[1]}, {u[1], u[2]}, {u[0], u[1]}, {u[1], u[2]}};
        for (int f = 0; f < 6; ++f) {
            faceDerivative(f, fu[f][0], fu[f][1], g2);
            for (int c = 0; c < 3; ++c) { fd[f][c][0] = 2.0 * g2[c][0]; fd[f][c][1] = 2.0 * g2[c][1]; }
        }
        const double x1 = 0.5 * (u[0] + 1.0), x2 = 0.5 * (u[1] + 1.0), x3 = 0.5 * (u[2] + 1.0);
        const double eps = 100.0 * 2.220446049250313e-16;
        for (int j = 0; j < 3; ++j) {
            const double c1 = corners[0][j], c2 = corners[1][j], c3 = corners[2][j], c4 = corners[3][j], c5 = corners[4][j], c6 = corners[5][j], c7 = corners[6][j], c8 = corners[7][j];
            double g1 = -face[5][j] + face[3][j] + fd[0][j][0] * (1.0 - x2) + fd[1][j][0] * x2 + fd[2][j][0] * (1.0 - x3) + fd[4][j][0] * x3;
            g1 = g1 - ed[0][j] * (1.0 - x2) * (1.0 - x3) - ed[2][j] * (1.0 - x2) * x3 - ed[4][j] * x2 * (1.0 - x3) - ed[6][j] * x2 * x3
                    + edge[8][j] * (1.0 - x3) + edge[11][j] * x3 - edge[9][j] * (1.0 - x3) - edge[10][j] * x3
                    + edge[3][j] * (1.0 - x2) + edge[7][j] * x2 - edge[1][j] * (1.0 - x2) - edge[5][j] * x2;
            g1 = g1 - c1 * (1.0 - x2) * (1.0 - x3) - c5 * (1.0 - x2) * x3 - c4 * x2 * (1.0 - x3) - c8 * x2 * x3
                    + c2 * (1.0 - x2) * (1.0 - x3) + c6 * (1.0 - x2) * x3 + c3 * x2 * (1.0 - x3) + c7 * x2 * x3;
            double g2_ = fd[5][j][0] * (1.0 - x1) + fd[3][j][0] * x1 - face[0][j] + face[1][j] + fd[2][j][1] * (1.0 - x3) + fd[4][j][1] * x3;
            g2_ = g2_ + edge[0][j] * (1.0 - x3) + edge[2][j] * x3 - edge[4][j] * (1.0 - x3) - edge[6][j] * x3
                      - ed[8][j] * (1.0 - x1) * (1.0 - x3) - ed[11][j] * (1.0 - x1) * x3 - ed[9][j] * (1.0 - x3) * x1 - ed[10][j] * x1 * x3
                      + edge[3][j] * (1.0 - x1) - edge[7][j] * (1.0 - x1) + edge[1][j] * x1 - edge[5][j] * x1;
            g2_ = g2_ - c1 * (1.0 - x1) * (1.0 - x3) - c5 * (1.0 - x1) * x3 + c4 * (1.0 - x1) * (1.0 - x3) + c8 * (1.0 - x1) * x3
                      - c2 * x1 * (1.0 - x3) - c6 * x1 * x3 + c3 * x1 * (1.0 - x3) + c7 * x1 * x3;
            double g3 = fd[5][j][1] * (1.0 - x1) + fd[3][j][1] * x1 + fd[0][j][1] * (1.0 - x2) + fd[1][j][1] * x2 - face[2][j] + face[4][j];
            g3 = g3 + edge[0][j] * (1.0 - x2) - edge[2][j] * (1.0 - x2) + edge[4][j] * x2 - edge[6][j] * x2
                    + edge[8][j] * (1.0 - x1) - edge[11][j] * (1.0 - x1) + edge[9][j] * x1 - edge[10][j] * x1
                    - ed[3][j] * (1.0 - x1) * (1.0 - x2) - ed[7][j] * (1.0 - x1) * x2 - ed[1][j] * x1 * (1.0 - x2) - ed[5][j] * x1 * x2;
            g3 = g3 - c1 * (1.0 - x1) * (1.0 - x2) + c5 * (1.0 - x1) * (1.0 - x2) - c4 * (1.0 - x1) * x2 + c8 * (1.0 - x1) * x2
                    - c2 * x1 * (1.0 - x2) + c6 * x1 * (1.0 - x2) - c3 * x1 * x2 + c7 * x1 * x2;
            gg[j][0] = 0.5 * g1; gg[j][1] = 0.5 * g2_; gg[j][2] = 0.5 * g3;
            for (int d = 0; d < 3; ++d) if (std::fabs(gg[j][d]) <= eps) gg[j][d] = 0.0;
        }
    }
    void facePoint(int f, double u, double v, double* p) const {
        const FacePatch& fp = m->patches[e][f];
        if (fp.nu == 2 && fp.nv == 2) {
            for (int c = 0; c < 3; ++c)
                p[c] = 0.25 * (fp.pts[0 * 3 + c] * (1 - u) * (1 - v) + fp.pts[1 * 3 + c] * (1 + u) * (1 - v) +
                               fp.pts[3 * 3 + c] * (1 + u) * (1 + v) + fp.pts[2 * 3 + c] * (1 - u) * (1 + v));
            return;
        }
        double li[32], lj[32];
        interpolatingPolynomialVector(u, fp.nu - 1, knots[f][0].data(), wb[f][0].data(), li);
        interpolatingPolynomialVector(v, fp.nv - 1, knots[f][1].data(), wb[f][1].data(), lj);
        p[0] = p[1] = p[2] = 0.0;
        for (int j = 0; j < fp.nv; ++j) for (int i = 0; i < fp.nu; ++i) {
            double l = li[i] * lj[j];
            for (int c = 0; c < 3; ++c) p[c] += fp.pts[(j * fp.nu + i) * 3 + c] * l;
        }
    }
    void at(const double* u, double* x) const {
        double xi[3] = {0.5 * (u[0] + 1.0), 0.5 * (u[1] + 1.0), 0.5 * (u[2] + 1.0)};
        if (m->isHex8[e]) {
            for (int j = 0; j < 3; ++j)
                x[j] = corners[0][j] * (1 - xi[0]) * (1 - xi[1]) * (1 - xi[2]) + corners[1][j] * xi[0] * (1 - xi[1]) * (1 - xi[2]) +
                       corners[2][j] * xi[0] * xi[1] * (1 - xi[2]) + corners[3][j] * (1 - xi[0]) * xi[1] * (1 - xi[2]) +
                       corners[4][j] * (1 - xi[0]) * (1 - xi[1]) * xi[2] + corners[5][j] * xi[0] * (1 - xi[1]) * xi[2] +
                       corners[6][j] * xi[0] * xi[1] * xi[2] + corners[7][j] * (1 - xi[0]) * xi[1] * xi[2];
            return;
        }
        double face[6][3], edge[12][3];
        facePoint(0, u[0], -1.0, edge[0]); facePoint(0, 1.0, u[2], edge[1]); facePoint(0, u[0], 1.0, edge[2]); facePoint(0, -1.0, u[2], edge[3]);
        facePoint(1, u[0], -1.0, edge[4]); facePoint(1, 1.0, u[2], edge[5]); facePoint(1, u[0], 1.0, edge[6]); facePoint(1, -1.0, u[2], edge[7]);
        facePoint(3, u[1], -1.0, edge[9]); facePoint(5, u[1], -1.0, edge[8]); facePoint(3, u[1], 1.0, edge[10]); facePoint(5, u[1], 1.0, edge[11]);
        facePoint(0, u[0], u[2], face[0]); facePoint(1, u[0], u[2], face[1]); facePoint(2, u[0], u[1], face[2]);
        facePoint(3, u[1], u[2], face[3]); facePoint(4, u[0], u[1], face[4]); facePoint(5, u[1], u[2], face[5]);
        for (int j = 0; j < 3; ++j) {
            double r = face[5][j] * (1 - xi[0]) + face[3][j] * xi[0] + face[0][j] * (1 - xi[1]) + face[1][j] * xi[1] + face[2][j] * (1 - xi[2]) + face[4][j] * xi[2];
            r = r - edge[0][j] * (1 - xi[1]) * (1 - xi[2]) - edge[2][j] * (1 - xi[1]) * xi[2] - edge[4][j] * xi[1] * (1 - xi[2]) - edge[6][j] * xi[1] * xi[2]
                  - edge[8][j] * (1 - xi[0]) * (1 - xi[2]) - edge[11][j] * (1 - xi[0]) * xi[2] - edge[9][j] * (1 - xi[2]) * xi[0] - edge[10][j] * xi[0] * xi[2]
                  - edge[3][j] * (1 - xi[0]) * (1 - xi[1]) - edge[7][j] * (1 - xi[0]) * xi[1] - edge[1][j] * xi[0] * (1 - xi[1]) - edge[5][j] * xi[0] * xi[1];
            r = r + corners[0][j] * (1 - xi[0]) * (1 - xi[1]) * (1 - xi[2]) + corners[4][j] * (1 - xi[0]) * (1 - xi[1]) * xi[2]
                  + corners[3][j] * (1 - xi[0]) * xi[1] * (1 - xi[2]) + corners[7][j] * (1 - xi[0]) * xi[1] * xi[2]
                  + corners[1][j] * xi[0] * (1 - xi[1]) * (1 - xi[2]) + corners[5][j] * xi[0] * (1 - xi[1]) * xi[2]
                  + corners[2][j] * xi[0] * xi[1] * (1 - xi[2]) + corners[6][j] * xi[0] * xi[1] * xi[2];
            x[j] = r;
        }
    }
    // gradient of the hex8 map: g[i][j] = d x_i / d xi_j
    void gradHex8(const double* u, double g[3][3]) const {
        double xi[3] = {0.5 * (u[0] + 1.0), 0.5 * (u[1] + 1.0), 0.5 * (u[2] + 1.0)};
        for (int i = 0; i < 3; ++i) {
            g[i][0] = 0.5 * (-corners[0][i] * (1 - xi[1]) * (1 - xi[2]) + corners[1][i] * (1 - xi[1]) * (1 - xi[2]) + corners[2][i] * xi[1] * (1 - xi[2]) - corners[3][i] * xi[1] * (1 - xi[2])
                             - corners[4][i] * (1 - xi[1]) * xi[2] + corners[5][i] * (1 - xi[1]) * xi[2] + corners[6][i] * xi[1] * xi[2] - corners[7][i] * xi[1] * xi[2]);
            g[i][1] = 0.5 * (-corners[0][i] * (1 - xi[0]) * (1 - xi[2]) - corners[1][i] * xi[0] * (1 - xi[2]) + corners[2][i] * xi[0] * (1 - xi[2]) + corners[3][i] * (1 - xi[0]) * (1 - xi[2])
                             - corners[4][i] * (1 - xi[0]) * xi[2] - corners[5][i] * xi[0] * xi[2] + corners[6][i] * xi[0] * xi[2] + corners[7][i] * (1 - xi[0]) * xi[2]);
            g[i][2] = 0.5 * (-corners[0][i] * (1 - xi[0]) * (1 - xi[1]) - corners[1][i] * xi[0] * (1 - xi[1]) - corners[2][i] * xi[0] * xi[1] - corners[3][i] * (1 - xi[0]) * xi[1]
                             + corners[4][i] * (1 - xi[0]) * (1 - xi[1]) + corners[5][i] * xi[0] * (1 - xi[1]) + corners[6][i] * xi[0] * xi[1] + corners[7][i] * (1 - xi[0]) * xi[1]);
        }
    }
};

// local element trace index of face node (a,b) on local face f -> (i,j,k) with the normal index left free
inline void faceNodeToElem(int f, int a, int b, int nrm, int& i, int& j, int& k) {
    int idx[3]; idx[axisMap[f][0]] = a; idx[axisMap[f][1]] = b; idx[faceNormalAxis[f]] = nrm;
    i = idx[0]; j = idx[1]; k = idx[2];
}

inline void buildGeometry(const HostMesh& m, int N, int nodeType, HostGeometry& g, bool referenceOrder = false) {
    g.N = N; g.n = N + 1; g.nodeType = nodeType; g.sp.construct(nodeType, N);
    const NodalStorage& sp = g.sp;
    const int n = g.n, n3 = n * n * n, n2 = n * n;
    const int nE = m.nElem();
    g.x.assign(3 * (size_t)n3 * nE, 0.0); g.jGradXi.assign(3 * (size_t)n3 * nE, 0.0); g.jGradEta.assign(3 * (size_t)n3 * nE, 0.0);
    g.jGradZeta.assign(3 * (size_t)n3 * nE, 0.0); g.jac.assign((size_t)n3 * nE, 0.0); g.invJac.assign((size_t)n3 * nE, 0.0); g.volume.assign(nE, 0.0);
#pragma omp parallel
    {
        std::vector<double> xC(3 * n3), gradx(9 * n3), cp(3 * n3), aux(9 * n3), Ja[3], JC(n3), tmp1(3 * n3), tmp2(3 * n3);
        for (int d = 0; d < 3; ++d) Ja[d].assign(3 * n3, 0.0);
        ElemMap map;
#pragma omp for schedule(static)
        for (int e = 0; e < nE; ++e) {
            map.init(m, e);
            auto I = [&](int i, int j, int k) { return (k * n + j) * n + i; };
            // node coordinates
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                double u[3] = {sp.x[i], sp.x[j], sp.x[k]};
                map.at(u, &g.x[3 * ((size_t)e * n3 + I(i, j, k))]);
            }
            // mapping and its gradient on the Chebyshev-Gauss-Lobatto grid
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                double u[3] = {sp.xCGL[i], sp.xCGL[j], sp.xCGL[k]};
                map.at(u, &xC[3 * I(i, j, k)]);
            }
            if (m.isHex8[e]) {
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                    double u[3] = {sp.xCGL[i], sp.xCGL[j], sp.xCGL[k]}, gg[3][3];
                    map.gradHex8(u, gg);
                    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) gradx[9 * I(i, j, k) + 3 * a + b] = gg[a][b];
                }
            } else {
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                    double u[3] = {sp.xCGL[i], sp.xCGL[j], sp.xCGL[k]}, gg[3][3];
                    map.gradGeneral(u, gg);
                    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) gradx[9 * I(i, j, k) + 3 * a + b] = gg[a][b];
                }
            }
            // curl form: component c of Ja^d = -1/2 [curl( X_l grad X_m - X_m grad X_l )]_d, (c,m,l) cyclic
            for (int c = 0; c < 3; ++c) {
                const int l_ = (c + 2) % 3, m_ = (c + 1) % 3;   // c=0: X3 grad X2 - X2 grad X3
                for (int q = 0; q < n3; ++q) for (int d = 0; d < 3; ++d)
                    cp[3 * q + d] = xC[3 * q + l_] * gradx[9 * q + 3 * m_ + d] - xC[3 * q + m_] * gradx[9 * q + 3 * l_ + d];
                std::fill(aux.begin(), aux.end(), 0.0);
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                    double* a = &aux[9 * I(i, j, k)];  // a[3*comp + dir]
                    for (int l = 0; l < n; ++l) for (int d = 0; d < 3; ++d) {
                        a[3 * d + 0] += cp[3 * I(l, j, k) + d] * sp.DCGL[i * n + l];
                        a[3 * d + 1] += cp[3 * I(i, l, k) + d] * sp.DCGL[j * n + l];
                        a[3 * d + 2] += cp[3 * I(i, j, l) + d] * sp.DCGL[k * n + l];
                    }
                }
                for (int q = 0; q < n3; ++q) {
                    const double* a = &aux[9 * q];
                    double J1 = a[3 * 2 + 1] - a[3 * 1 + 2], J2 = a[3 * 0 + 2] - a[3 * 2 + 0], J3 = a[3 * 1 + 0] - a[3 * 0 + 1];
                    Ja[0][3 * q + c] = -0.5 * J1; Ja[1][3 * q + c] = -0.5 * J2; Ja[2][3 * q + c] = -0.5 * J3;
                }
            }
            for (int q = 0; q < n3; ++q) {
                const double* G = &gradx[9 * q];   // G[3*i + j] = dx_i/dxi_j ; a_j = column j
                double a1[3] = {G[0], G[3], G[6]}, a2[3] = {G[1], G[4], G[7]}, a3[3] = {G[2], G[5], G[8]};
                JC[q] = a1[0] * (a2[1] * a3[2] - a2[2] * a3[1]) + a1[1] * (a2[2] * a3[0] - a2[0] * a3[2]) + a1[2] * (a2[0] * a3[1] - a2[1] * a3[0]);
            }
            // back to the solution nodes (sum-factorised TCheb2Gauss in xi, eta, zeta)
            auto interp3 = [&](const double* src, int nc, double* dst) {
                if (referenceOrder) {   // MappedGeometry.f90:336-361: the full triple sum, l fastest, products left to right
                    for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                        double acc[3] = {0.0, 0.0, 0.0};
                        for (int nn = 0; nn < n; ++nn) for (int mm = 0; mm < n; ++mm) for (int l = 0; l < n; ++l)
                            for (int c = 0; c < nc; ++c)
                                acc[c] = acc[c] + src[nc * I(l, mm, nn) + c] * sp.TCheb2Gauss[i * n + l] * sp.TCheb2Gauss[j * n + mm] * sp.TCheb2Gauss[k * n + nn];
                        for (int c = 0; c < nc; ++c) dst[nc * I(i, j, k) + c] = acc[c];
                    }
                    return;
                }
                double* t1 = tmp1.data(); double* t2 = tmp2.data();
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) for (int c = 0; c < nc; ++c) {
                    double s = 0; for (int l = 0; l < n; ++l) s += src[nc * I(l, j, k) + c] * sp.TCheb2Gauss[i * n + l];
                    t1[nc * I(i, j, k) + c] = s;
                }
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) for (int c = 0; c < nc; ++c) {
                    double s = 0; for (int l = 0; l < n; ++l) s += t1[nc * I(i, l, k) + c] * sp.TCheb2Gauss[j * n + l];
                    t2[nc * I(i, j, k) + c] = s;
                }
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) for (int c = 0; c < nc; ++c) {
                    double s = 0; for (int l = 0; l < n; ++l) s += t2[nc * I(i, j, l) + c] * sp.TCheb2Gauss[k * n + l];
                    dst[nc * I(i, j, k) + c] = s;
                }
            };
            interp3(Ja[0].data(), 3, &g.jGradXi[3 * (size_t)e * n3]);
            interp3(Ja[1].data(), 3, &g.jGradEta[3 * (size_t)e * n3]);
            interp3(Ja[2].data(), 3, &g.jGradZeta[3 * (size_t)e * n3]);
            interp3(JC.data(), 1, &g.jac[(size_t)e * n3]);
            double vol = 0.0;
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                size_t q = (size_t)e * n3 + I(i, j, k);
                g.invJac[q] = 1.0 / g.jac[q];
                vol += sp.w[i] * sp.w[j] * sp.w[k] * g.jac[q];
            }
            g.volume[e] = vol;
        }
    }
    // ---- faces (geometry from the LEFT element; MPI faces are rebuilt per partition by the caller)
    const int nF = m.nFaces;
    g.fx.assign(3 * (size_t)n2 * nF, 0.0); g.fnormal.assign(3 * (size_t)n2 * nF, 0.0); g.ft1.assign(3 * (size_t)n2 * nF, 0.0);
    g.ft2.assign(3 * (size_t)n2 * nF, 0.0); g.fjac.assign((size_t)n2 * nF, 0.0); g.fsurface.assign(nF, 0.0);
#pragma omp parallel
    {
        ElemMap map;
#pragma omp for schedule(static)
        for (int f = 0; f < nF; ++f) {
            int side = 0;
            if (m.faceElem[2 * f] < 0) side = 1;
            const int e = m.faceElem[2 * f + side], lf = m.faceElemSide[2 * f + side];
            const int rot = side == 0 ? 0 : m.faceRot[f];
            map.init(m, e);
            const double* v = &sp.v[faceNormalEnd[lf] * n];
            const double* Jd = (faceNormalAxis[lf] == 0 ? g.jGradXi.data() : faceNormalAxis[lf] == 1 ? g.jGradEta.data() : g.jGradZeta.data()) + 3 * (size_t)e * n3;
            std::vector<double> dS(3 * n2, 0.0);
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
                int idx[3] = {i, j, k};
                int a = idx[axisMap[lf][0]], b = idx[axisMap[lf][1]], nr = idx[faceNormalAxis[lf]];
                for (int c = 0; c < 3; ++c) dS[3 * (b * n + a) + c] += Jd[3 * ((k * n + j) * n + i) + c] * v[nr];
            }
            double sgn = faceNormalEnd[lf] ? 1.0 : -1.0;
            if (side == 1) sgn = -sgn;
            for (int b = 0; b < n; ++b) for (int a = 0; a < n; ++a) {
                int aa = a, bb = b;
                if (rot != 0) leftIndexes2Right(a, b, N, N, rot, aa, bb);
                const size_t q = (size_t)f * n2 + b * n + a;
                double nv[3] = {sgn * dS[3 * (bb * n + aa)], sgn * dS[3 * (bb * n + aa) + 1], sgn * dS[3 * (bb * n + aa) + 2]};
                double nrm = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
                g.fjac[q] = nrm;
                for (int c = 0; c < 3; ++c) g.fnormal[3 * q + c] = nv[c] / nrm;
                // face coordinates (coordRotation: face node -> element face coordinates)
                double xi = sp.x[a], eta = sp.x[b], xr, er;
                switch (rot) {   // MeshTypes.f90 coordRotation
                    case 0: xr = xi; er = eta; break;
                    case 1: xr = -eta; er = xi; break;
                    case 2: xr = -xi; er = -eta; break;
                    case 3: xr = eta; er = -xi; break;
                    case 4: xr = eta; er = xi; break;
                    case 5: xr = -xi; er = eta; break;
                    case 6: xr = -eta; er = -xi; break;
                    default: xr = xi; er = -eta; break;
                }
                double u[3]; u[axisMap[lf][0]] = xr; u[axisMap[lf][1]] = er; u[faceNormalAxis[lf]] = faceNormalEnd[lf] ? 1.0 : -1.0;
                map.at(u, &g.fx[3 * q]);
            }
            double surf = 0.0;
            for (int b = 0; b < n; ++b) for (int a = 0; a < n; ++a) {
                const size_t q = (size_t)f * n2 + b * n + a;
                double t1[3] = {0, 0, 0};
                for (int l = 0; l < n; ++l) for (int c = 0; c < 3; ++c) t1[c] += sp.D[a * n + l] * g.fx[3 * ((size_t)f * n2 + b * n + l) + c];
                const double* nh = &g.fnormal[3 * q];
                double dot = t1[0] * nh[0] + t1[1] * nh[1] + t1[2] * nh[2];
                for (int c = 0; c < 3; ++c) t1[c] -= dot * nh[c];
                double nt = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
                for (int c = 0; c < 3; ++c) { t1[c] /= nt; g.ft1[3 * q + c] = t1[c]; }
                g.ft2[3 * q + 0] = nh[1] * t1[2] - nh[2] * t1[1];
                g.ft2[3 * q + 1] = nh[2] * t1[0] - nh[0] * t1[2];
                g.ft2[3 * q + 2] = nh[0] * t1[1] - nh[1] * t1[0];
                surf += sp.w[a] * sp.w[b] * g.fjac[q];
            }
            g.fsurface[f] = surf;
        }
    }
    computeFaceMinimumDistance(m, g);
}

// Faces' minimum orthogonal distance estimate (HexMesh.f90:3016-3041): h = min over the adjacent elements of min(J) / max(J_f).
// A face on a partition cut sees its local element only; the reference then takes the minimum with the neighbour's value
// (CommunicateMPIFaceMinimumDistance, HexMesh.f90:3059-3145) -- partitions made with inherit_geometry copy the global value.
inline void computeFaceMinimumDistance(const HostMesh& m, HostGeometry& g) {
    const size_t n2 = (size_t)g.n * g.n, n3 = n2 * g.n;
    std::vector<double> minJ(m.nElem());
    for (int e = 0; e < m.nElem(); ++e) minJ[e] = *std::min_element(g.jac.begin() + e * n3, g.jac.begin() + (e + 1) * n3);
    g.fh.assign(m.nFaces, 0.0);
    for (int f = 0; f < m.nFaces; ++f) {
        const int e1 = m.faceElem[2 * f], e2 = m.faceElem[2 * f + 1];
        double num;
        if (e1 >= 0 && e2 >= 0) num = std::min(minJ[e1], minJ[e2]);
        else num = minJ[std::max(e1, e2)];
        g.fh[f] = num / *std::max_element(g.fjac.begin() + f * n2, g.fjac.begin() + (f + 1) * n2);
    }
}

// Coordinates of the nodes of this mesh's no-slip wall faces (the local contribution to GatherAllWallCoordinates,
// HexMesh.f90:5696-5780)
inline std::vector<double> wallCoordinates(const HostMesh& m, const HostGeometry& g) {
    const int n2 = g.n * g.n;
    std::vector<double> Xw;
    for (int f = 0; f < m.nFaces; ++f) {
        if (m.faceType[f] != HMESH_BOUNDARY || m.faceZone[f] < 0 || m.bcs[m.faceZone[f]].type != "noslipwall") continue;
        Xw.insert(Xw.end(), &g.fx[3 * (size_t)f * n2], &g.fx[3 * (size_t)f * n2] + 3 * n2);
    }
    return Xw;
}

// HexMesh_ComputeWallDistances (HexMesh.f90:5594-5692): distance of every element and face node to the nearest no-slip wall
// node.  Xw holds the wall nodes of ALL partitions (the reference gathers them across ranks first); a partition that scanned
// its own wall faces only would make LS = min(Cs delta, 0.4 dWall) depend on the partition.
inline void computeWallDistances(const HostMesh& m, HostGeometry& g, const std::vector<double>& Xw) {
    const int n = g.n, n2 = n * n, n3 = n2 * n;
    const size_t nW = Xw.size() / 3;
    auto dist = [&](const double* xP) {
        double mn = 1.7976931348623157e308;
        for (size_t q = 0; q < nW; ++q) {
            const double d0 = xP[0] - Xw[3 * q], d1 = xP[1] - Xw[3 * q + 1], d2 = xP[2] - Xw[3 * q + 2];
            const double cur = d0 * d0 + d1 * d1 + d2 * d2;
            mn = std::fmin(mn, cur);
        }
        return std::sqrt(mn);
    };
    g.dWall.resize((size_t)m.nElem() * n3); g.fdWall.resize((size_t)m.nFaces * n2);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)g.dWall.size(); ++q) g.dWall[q] = dist(&g.x[3 * q]);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < (long long)g.fdWall.size(); ++q) g.fdWall[q] = dist(&g.fx[3 * q]);
}
inline void computeWallDistances(const HostMesh& m, HostGeometry& g) { computeWallDistances(m, g, wallCoordinates(m, g)); }

}  // namespace h3d
