// ======================================================================================================
//  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Never loaded by horses3d_b200/, bench.py or the C ABI.
//
//  Host-loop backend of the p-nonconforming device path.  horses3d_b200/csrc/h3d_mixed.cuh writes every kernel of that path
//  as a functor over the thread index and its orchestration (allocation sizes, tables, launch sequence, reductions) as a
//  template over a backend.  libh3dgpu.so instantiates it with the CUDA backend; this file instantiates the SAME functors and
//  the SAME orchestration with a backend whose "launch" is a loop over the thread index (in a random order, to expose any
//  dependence between threads of one launch), and exports them under the prefix emu_ with the signatures of include/h3d_gpu.h.
//  tests/test_mixed_emu.py compares it with the oracle where no GPU is present.  What it cannot see: CUDA launch
//  configuration and anything specific to nvcc's code generation -- the -m gpu tests cover those on the device.
// ======================================================================================================
#include <math.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <random>
#include <string>
#include <utility>
#include <vector>
#define __noinline__      // a CUDA function qualifier of h3d_physics.cuh; empty so that no standard header can trip over it
#include "h3d_mixed.cuh"

using namespace h3d;

#include <condition_variable>
#include <mutex>

namespace {
// Ranks of one emulated run live in one process, one thread each (the Python test calls the entry points from threads; ctypes
// releases the interpreter lock).  exchange / allreduce meet at a barrier and copy through the posted pointers.
struct World {
    int nranks; std::mutex mu; std::condition_variable cv; int arrived = 0; long long generation = 0;
    std::vector<const double*> sendBuf; std::vector<std::vector<int>> nbrRanks; std::vector<std::vector<long long>> nbrOff;
    std::vector<const double*> scal;
    explicit World(int n) : nranks(n), sendBuf(n), nbrRanks(n), nbrOff(n), scal(n) {}
    bool aborted = false;      // a rank failed: the barriers open so that the other ranks run to the end and the test can report
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return;
        const long long gen = generation;
        if (++arrived == nranks) { arrived = 0; ++generation; cv.notify_all(); } else cv.wait(lk, [&] { return generation != gen || aborted; });
    }
    void abort() { std::unique_lock<std::mutex> lk(mu); aborted = true; cv.notify_all(); }
};
struct HostBackend {
    World* world = nullptr; int rank = 0;
    void exchange(const double* send, double* recv, int nNbr, const int* ranks, const long long* off, const long long* cnt) {
        world->sendBuf[rank] = send; world->nbrRanks[rank].assign(ranks, ranks + nNbr); world->nbrOff[rank].assign(off, off + nNbr);
        world->barrier();
        for (int b = 0; b < nNbr; ++b) {   // my message from neighbour r = the part of r's send buffer addressed to me
            const int r = ranks[b]; long long roff = -1;
            for (size_t q = 0; q < world->nbrRanks[r].size(); ++q) if (world->nbrRanks[r][q] == rank) roff = world->nbrOff[r][q];
            if (roff < 0 || world->aborted) continue;
            std::memcpy(recv + off[b], world->sendBuf[r] + roff, (size_t)cnt[b] * sizeof(double));
        }
        world->barrier();
    }
    void allreduce(double* v, int n, int op) {
        world->scal[rank] = v;
        world->barrier();
        std::vector<double> out(v, v + n);
        if (!world->aborted) for (int q = 0; q < n; ++q) {
            double acc = world->scal[0][q];
            for (int r = 1; r < world->nranks; ++r) { const double x = world->scal[r][q]; acc = op == 0 ? std::fmax(acc, x) : (op == 1 ? std::fmin(acc, x) : acc + x); }
            out[q] = acc;
        }
        world->barrier();
        std::memcpy(v, out.data(), n * sizeof(double));
        world->barrier();
    }
    std::vector<void*>* allocs = nullptr;
    template <class T> T* alloc(size_t count) { void* q = std::malloc(std::max<size_t>(count, 1) * sizeof(T)); allocs->push_back(q); return (T*)q; }
    template <class T> void zero(T* dst, size_t count) { std::memset(dst, 0, count * sizeof(T)); }
    template <class T> void upload(T* dst, const T* src, size_t count) { std::memcpy(dst, src, count * sizeof(T)); }
    template <class T> void download(T* dst, const T* src, size_t count) { std::memcpy(dst, src, count * sizeof(T)); }
    template <class F> void launch(const F& f, long long count) {
        // a permutation of the thread indices: the result must not depend on the order in which the threads of a launch run
        std::vector<long long> order(count);
        std::iota(order.begin(), order.end(), 0LL);
        std::mt19937_64 rng(12345u + (unsigned)count);
        std::shuffle(order.begin(), order.end(), rng);
#pragma omp parallel for schedule(static)
        for (long long t = 0; t < count; ++t) f(order[t]);
    }
    const char* error() { return nullptr; }
};
struct Emu {
    std::vector<void*> allocs;
    H3dPhysics physics{}; Phys ph{}; bool havePhysics = false;
    MixedSolver<HostBackend>* mx = nullptr;
    std::string err;
    explicit Emu(World* w = nullptr, int rank = 0) {
        HostBackend be; be.allocs = &allocs; be.world = w; be.rank = rank;
        mx = new MixedSolver<HostBackend>(be);
        if (w) mx->nranks = w->nranks;
    }
    ~Emu() { delete mx; for (void* p : allocs) std::free(p); }
};
int done(Emu* h, int rc) { if (rc) h->err = h->mx->err; return rc; }
}  // namespace

extern "C" {
int emu_create_handle(void** out, int, int, int, const void*) { *out = new Emu(); return 0; }
void* emu_create() { return new Emu(); }
void* emu_world_create(int nranks) { return new World(nranks); }
void emu_world_destroy(void* w) { delete (World*)w; }
void emu_world_abort(void* w) { ((World*)w)->abort(); }
void* emu_create_rank(void* world, int rank) { return new Emu((World*)world, rank); }
int emu_set_halo(void* p, int nNbr, const int* ranks, const int* counts, const int* faces, const int* sides) { Emu* h = (Emu*)p; return done(h, h->mx->setHalo(nNbr, ranks, counts, faces, sides)); }
void emu_destroy(void* p) { delete (Emu*)p; }
const char* emu_last_error(void* p) { return ((Emu*)p)->err.c_str(); }
int emu_set_physics(void* p, const H3dPhysics* ph) {
    Emu* h = (Emu*)p;
    if (physFromH3dPhysics(ph, h->ph, h->err)) return 1;
    h->physics = *ph; h->havePhysics = true;
    return 0;
}
int emu_set_basis(void* p, int N, int nodeType, const double* x, const double* w, const double* D, const double* hatD, const double* sharpD, const double* v, const double* b) {
    ((Emu*)p)->mx->setBasis(N, nodeType, x, w, D, hatD, sharpD, v, b); return 0;
}
int emu_set_interpolation(void* p, int No, int Nd, const double* T) { Emu* h = (Emu*)p; return done(h, h->mx->setInterpolation(No, Nd, T)); }
int emu_set_mesh_p(void* p, int nElem, int nFace, const int* elemOrder, const int* faceOrder, const int* elemFace, const int* elemFaceSide, const int* faceElem,
                   const int* faceElemSide, const int* faceRot, const int* faceType, const int* faceZone, const double* jGradXi, const double* jGradEta,
                   const double* jGradZeta, const double* jacobian, const double*, const double* volume, const double* faceNormal, const double* faceT1,
                   const double* faceT2, const double* faceJacobian, const double*, const double* faceSurface) {
    Emu* h = (Emu*)p;
    if (!h->havePhysics) { h->err = "set_physics must precede set_mesh_p"; return 1; }
    h->mx->ph = h->ph;
    return done(h, h->mx->setMesh(h->physics, nElem, nFace, elemOrder, faceOrder, elemFace, elemFaceSide, faceElem, faceElemSide, faceRot, faceType, faceZone, jGradXi,
                                   jGradEta, jGradZeta, jacobian, volume, faceNormal, faceT1, faceT2, faceJacobian, faceSurface));
}
int emu_set_wall_distance(void* p, const double* a, const double* b) { Emu* h = (Emu*)p; return done(h, h->mx->setWallDistance(a, b)); }
int emu_set_face_h(void* p, const double* fH) { Emu* h = (Emu*)p; return done(h, h->mx->setFaceH(h->physics, fH)); }
int emu_set_boundary_conditions(void* p, int nZones, const int* bcType, const double* bcParams) { Emu* h = (Emu*)p; return done(h, h->mx->setBoundaryConditions(nZones, bcType, bcParams)); }
int emu_upload_Q(void* p, const double* Q) { Emu* h = (Emu*)p; return done(h, h->mx->uploadQ(Q)); }
int emu_download(void* p, double* Q, double* QDot, double* Ux, double* Uy, double* Uz) { Emu* h = (Emu*)p; return done(h, h->mx->download(Q, QDot, Ux, Uy, Uz)); }
int emu_set_source(void* p, const double* S) { Emu* h = (Emu*)p; return done(h, h->mx->setSource(S)); }
int emu_compute_time_derivative(void* p, double) { Emu* h = (Emu*)p; return done(h, h->mx->residual(h->physics, MxRk{0, 0.0, 0.0, 0.0, 0})); }
int emu_rk_step(void* p, int scheme, double, double dt, int ctd) { Emu* h = (Emu*)p; return done(h, h->mx->rkStep(h->physics, scheme, dt, ctd)); }
int emu_rk_stage(void* p, int scheme, int stage, double, double dt) { Emu* h = (Emu*)p; return done(h, h->mx->rkStage(h->physics, scheme, stage, dt)); }
int emu_max_residuals(void* p, double* out) { Emu* h = (Emu*)p; int nan = 0; return done(h, h->mx->maxResiduals(out, &nan)); }
int emu_has_nan(void* p, int* flag) { Emu* h = (Emu*)p; double r[5]; return done(h, h->mx->maxResiduals(r, flag)); }
int emu_max_timestep(void* p, double cfl, double dcfl, double* a, double* b) { Emu* h = (Emu*)p; return done(h, h->mx->maxTimestep(cfl, dcfl, a, b)); }
int emu_volume_integral(void* p, int kind, double* val) { Emu* h = (Emu*)p; return done(h, h->mx->volumeIntegral(h->physics, kind, val)); }
int emu_surface_integral(void* p, int zone, int kind, double* out) { Emu* h = (Emu*)p; return done(h, h->mx->surfaceIntegral(h->physics, zone, kind, out)); }
int emu_probe(void* p, int n, const int* elem, const int* var, const double* lx, const double* ly, const double* lz, double* values) {
    Emu* h = (Emu*)p; return done(h, h->mx->probe(n, elem, var, lx, ly, lz, values));
}
int emu_enable_limiter(void* p, int enabled, double minimum) { Emu* h = (Emu*)p; return done(h, h->mx->enableLimiter(enabled, minimum)); }
int emu_statistics_update(void* p, int reset) { Emu* h = (Emu*)p; return done(h, h->mx->statisticsUpdate(h->physics, reset)); }
int emu_statistics_download(void* p, double* data, int* nVars, int* nSamples) { Emu* h = (Emu*)p; return done(h, h->mx->statisticsDownload(data, nVars, nSamples)); }
int emu_snapshot_begin(void* p) { Emu* h = (Emu*)p; return done(h, h->mx->snapshotBegin()); }
int emu_snapshot_end(void* p, double* Q) { Emu* h = (Emu*)p; return done(h, h->mx->snapshotEnd(Q)); }
long long emu_kernel_launches(void* p) { return ((Emu*)p)->mx->launches; }
}
