// Host-side 1-D nodal operators (plays the role of the reference's NodalStorage_t on the driver side).
//
// Reference behaviour followed (paths relative to /root/reference/Solver/src/libs):
//   spectral/LegendreAlgorithms.f90:137-212   Gauss nodes/weights (Newton, <=10 its, tol 4 eps)
//   spectral/LegendreAlgorithms.f90:275-358   Gauss-Lobatto nodes/weights
//   spectral/InterpolationAndDerivatives.f90:109-159,230-255,413-458,799-828
//   spectral/NodalStorageClass.f90:201-307    D, hatD, sharpD, v, b, xCGL, DCGL, TCheb2Gauss
//
// All matrices are row-major: M[i*n + l] == M(i,l) of the reference.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

namespace h3d {

constexpr int GAUSS = 1;
constexpr int GAUSSLOBATTO = 2;
constexpr double PI_RP = 3.141592653589793238462643;

inline bool almostEqual(double a, double b) {
    const double tol = 2.0 * 2.220446049250313e-16;
    if (a == 0.0 || b == 0.0) return std::fabs(a - b) <= tol;
    return std::fabs(b - a) <= tol * std::fmax(std::fabs(a), std::fabs(b));
}

inline void legendrePolyAndDerivative(int N, double x, double& L, double& dL) {
    L = 0.0; dL = 0.0;
    if (N == 0) { L = 1.0; dL = 0.0; return; }
    if (N == 1) { L = x; dL = 1.0; return; }
    double Lm2 = 1.0, dLm2 = 0.0, Lm1 = x, dLm1 = 1.0;
    for (int k = 2; k <= N; ++k) {
        L = ((2 * k - 1) * x * Lm1 - (k - 1) * Lm2) / k;
        dL = dLm2 + (2 * k - 1) * Lm1;
        Lm2 = Lm1; Lm1 = L; dLm2 = dLm1; dLm1 = dL;
    }
}

inline void gaussNodes(int N, std::vector<double>& x, std::vector<double>& w) {
    x.assign(N + 1, 0.0); w.assign(N + 1, 0.0);
    const double tol = 4.0 * 2.220446049250313e-16;
    if (N == 0) { x[0] = 0.0; w[0] = 2.0; return; }
    if (N == 1) { x[0] = -std::sqrt(1.0 / 3.0); w[0] = 1.0; x[1] = -x[0]; w[1] = w[0]; }
    else {
        for (int j = 0; j < (N + 1) / 2; ++j) {
            double xj = -std::cos((2 * j + 1) * PI_RP / (2 * N + 2)), L, dL;
            for (int k = 0; k <= 10; ++k) {
                legendrePolyAndDerivative(N + 1, xj, L, dL);
                double delta = -L / dL;
                xj += delta;
                if (std::fabs(delta) <= tol * std::fabs(xj)) break;
            }
            legendrePolyAndDerivative(N + 1, xj, L, dL);
            x[j] = xj; w[j] = 2.0 / ((1.0 - xj * xj) * dL * dL);
            x[N - j] = -xj; w[N - j] = w[j];
        }
    }
    if (N % 2 == 0) {
        double L = 0, dL = 0; legendrePolyAndDerivative(N + 1, 0.0, L, dL);
        x[N / 2] = 0.0; w[N / 2] = 2.0 / (dL * dL);
    }
}

inline void qAndL(int N, double x, double& Q, double& dQ, double& LN) {
    double Lm2 = 1.0, dLm2 = 0.0, Lm1 = x, dLm1 = 1.0, Lk = 0, dLk = 0;
    for (int k = 2; k <= N; ++k) {
        Lk = ((2 * k - 1) * x * Lm1 - (k - 1) * Lm2) / k;
        dLk = dLm2 + (2 * k - 1) * Lm1;
        Lm2 = Lm1; Lm1 = Lk; dLm2 = dLm1; dLm1 = dLk;
    }
    int k = N + 1;
    Lk = ((2 * k - 1) * x * Lm1 - (k - 1) * Lm2) / k;
    dLk = dLm2 + (2 * k - 1) * Lm1;
    Q = Lk - Lm2; dQ = dLk - dLm2; LN = Lm1;
}

inline void lobattoNodes(int N, std::vector<double>& x, std::vector<double>& w) {
    x.assign(N + 1, 0.0); w.assign(N + 1, 0.0);
    const double tol = 4.0 * 2.220446049250313e-16;
    if (N == 0) { x[0] = 0.0; w[0] = 2.0; return; }   // NodalStorageClass.f90:219-221
    if (N == 1) { x[0] = -1.0; w[0] = 1.0; x[1] = 1.0; w[1] = 1.0; return; }
    x[0] = -1.0; w[0] = 2.0 / (N * (N + 1)); x[N] = 1.0; w[N] = w[0];
    for (int j = 1; j < (N + 1) / 2; ++j) {
        double xj = -std::cos((j + 0.25) * PI_RP / N - 3.0 / (8 * N * PI_RP * (j + 0.25))), Q, dQ, LN;
        for (int k = 0; k <= 10; ++k) {
            qAndL(N, xj, Q, dQ, LN);
            double delta = -Q / dQ;
            xj += delta;
            if (std::fabs(delta) <= tol * std::fabs(xj)) break;
        }
        qAndL(N, xj, Q, dQ, LN);
        x[j] = xj; w[j] = 2.0 / (N * (N + 1) * LN * LN);
        x[N - j] = -xj; w[N - j] = w[j];
    }
    if (N % 2 == 0) {
        double L = 0, dL = 0; legendrePolyAndDerivative(N, 0.0, L, dL);
        x[N / 2] = 0.0; w[N / 2] = 2.0 / (N * (N + 1) * L * L);
    }
}

inline void barycentricWeights(int N, const double* x, double* w) {
    for (int j = 0; j <= N; ++j) w[j] = 1.0;
    for (int j = 1; j <= N; ++j)
        for (int k = 0; k < j; ++k) { w[k] *= (x[k] - x[j]); w[j] *= (x[j] - x[k]); }
    for (int j = 0; j <= N; ++j) w[j] = 1.0 / w[j];
}

inline void interpolatingPolynomialVector(double x, int N, const double* nodes, const double* wb, double* p) {
    bool match = false;
    for (int j = 0; j <= N; ++j) { p[j] = 0.0; if (almostEqual(x, nodes[j])) { p[j] = 1.0; match = true; } }
    if (match) return;
    double d = 0.0;
    for (int j = 0; j <= N; ++j) { double t = wb[j] / (x - nodes[j]); p[j] = t; d += t; }
    for (int j = 0; j <= N; ++j) p[j] /= d;
}

// D(i,j), row-major
inline void polynomialDerivativeMatrix(int N, const double* nodes, double* D) {
    std::vector<double> wb(N + 1);
    barycentricWeights(N, nodes, wb.data());
    const int n = N + 1;
    for (int i = 0; i <= N; ++i) {
        D[i * n + i] = 0.0;
        for (int j = 0; j <= N; ++j) if (j != i) {
            D[i * n + j] = wb[j] / (wb[i] * (nodes[i] - nodes[j]));
            D[i * n + i] -= D[i * n + j];
        }
    }
}

// T(k,j): new node k <- old node j ; T is (M+1) x (N+1) row-major
inline void polynomialInterpolationMatrix(int N, int M, const double* oldN, const double* wb, const double* newN, double* T) {
    for (int k = 0; k <= M; ++k) interpolatingPolynomialVector(newN[k], N, oldN, wb, T + (size_t)k * (N + 1));
}

struct NodalStorage {
    int N = -1, nodes = GAUSS, n = 0;
    std::vector<double> x, w, wb, v, b, D, hatD, sharpD, xCGL, wbCGL, DCGL, TCheb2Gauss;
    // v, b: [side*n + i], side 0 = LEFT/FRONT/BOTTOM (-1), side 1 = RIGHT/BACK/TOP (+1)
    void construct(int nodeType, int N_) {
        N = N_; nodes = nodeType; n = N + 1;
        if (nodeType == GAUSS) gaussNodes(N, x, w);
        else if (nodeType == GAUSSLOBATTO) lobattoNodes(N, x, w);
        else throw std::runtime_error("Undefined nodes choice");
        D.assign(n * n, 0.0); hatD.assign(n * n, 0.0); sharpD.assign(n * n, 0.0);
        polynomialDerivativeMatrix(N, x.data(), D.data());
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i)
            hatD[i * n + j] = D[j * n + i] * w[j] / w[i];
        if (nodeType == GAUSSLOBATTO && N != 0) {
            for (int i = 0; i < n * n; ++i) sharpD[i] = 2.0 * D[i];
            sharpD[0] = 2.0 * D[0] + 1.0 / w[0];
            sharpD[N * n + N] = 2.0 * D[N * n + N] - 1.0 / w[N];
        }
        wb.assign(n, 0.0); barycentricWeights(N, x.data(), wb.data());
        v.assign(2 * n, 0.0); b.assign(2 * n, 0.0);
        interpolatingPolynomialVector(1.0, N, x.data(), wb.data(), v.data() + n);
        interpolatingPolynomialVector(-1.0, N, x.data(), wb.data(), v.data());
        for (int s = 0; s < 2; ++s) for (int i = 0; i < n; ++i) b[s * n + i] = v[s * n + i] / w[i];
        xCGL.assign(n, 0.0);
        if (N != 0) for (int i = 0; i <= N; ++i) xCGL[i] = -std::cos(1.0 * i * PI_RP / N);
        wbCGL.assign(n, 0.0); barycentricWeights(N, xCGL.data(), wbCGL.data());
        DCGL.assign(n * n, 0.0); polynomialDerivativeMatrix(N, xCGL.data(), DCGL.data());
        TCheb2Gauss.assign(n * n, 0.0);
        polynomialInterpolationMatrix(N, N, xCGL.data(), wbCGL.data(), x.data(), TCheb2Gauss.data());
    }
};

}  // namespace h3d
