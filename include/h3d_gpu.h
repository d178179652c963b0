/* h3d_gpu.h -- C ABI of libh3dgpu.so, the B200 (sm_100a) replacement for HORSES3D's explicit
 * compressible Navier-Stokes residual and low-storage Runge-Kutta step.
 *
 * Every entry point cites the reference interface it replaces (paths relative to
 * /root/reference/Solver/src).  A Fortran driver binds these through ISO_C_BINDING (see
 * INTEGRATION.md); nothing here depends on torch, C++ or CUDA types.
 *
 * Conventions
 *   - all entry points return 0 on success, non-zero on error; h3d_last_error() gives the message
 *     (reference: errorMessage(STD_OUT) + error stop, libs/foundation/Includes.h:5).
 *   - indices are 0-based; "none" is -1 (the Fortran adapter subtracts 1 in its gather loops).
 *   - local faces 0..5 = EFRONT,EBACK,EBOTTOM,ERIGHT,ETOP,ELEFT (libs/mesh/MeshTypes.f90:17-18).
 *   - face sides 0/1 = left/right (FaceClass.f90:76).
 *   - element fields use the reference's own per-element order, elements concatenated in local ID
 *     order: A[e][k][j][i][c], c fastest (StorageClass.f90:63-67,423-429).  Face fields: A[f][j][i][c].
 *   - all floating point is double (RP, libs/foundation/SMConstants.f90:9-12).
 *   - one context <-> one rank <-> one GPU; calls are asynchronous on the context's CUDA streams and
 *     synchronise only where a scalar or a host buffer is returned.
 */
#ifndef H3D_GPU_H
#define H3D_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h3d_context* h3d_handle;

/* node sets (libs/spectral/NodalStorageClass.f90:16-17) */
#define H3D_GAUSS 1
#define H3D_GAUSSLOBATTO 2

/* inviscid discretization (NavierStokesSolver/SpatialDiscretization.f90:85-99) */
#define H3D_STANDARD_DG 0
#define H3D_SPLIT_DG 1

/* Riemann solvers (libs/physics/navierstokes/RiemannSolvers_NS.f90:120-206) */
#define H3D_RIEMANN_ROE 0
#define H3D_RIEMANN_LXF 1
#define H3D_RIEMANN_CENTRAL 2
#define H3D_RIEMANN_RUSANOV 3
#define H3D_RIEMANN_STDROE 4
#define H3D_RIEMANN_UDISS 5
#define H3D_RIEMANN_ROEPIKE 6
#define H3D_RIEMANN_LOWDISSROE 7
#define H3D_RIEMANN_MATRIXDISS 8

/* averaging / two-point flux (RiemannSolvers_NS.f90:233-285) */
#define H3D_AVG_STANDARD 0
#define H3D_AVG_KENNEDYGRUBER 1
#define H3D_AVG_PIROZZOLI 2
#define H3D_AVG_DUCROS 3
#define H3D_AVG_MORINISHI 4
#define H3D_AVG_ENTROPYCONS 5
#define H3D_AVG_CHANDRASEKAR 6

/* LES (libs/physics/common/LESModels.f90) */
#define H3D_LES_NONE 0
#define H3D_LES_SMAGORINSKY 1
#define H3D_LES_WALE 2            /* WALE_ComputeViscosity, LESModels.f90:358-435; intensity Cw (default 0.325)   */
#define H3D_LES_VREMAN 3          /* Vreman_ComputeViscosity, LESModels.f90:487-546; intensity C (default 0.07)    */

/* viscous discretization (libs/discretization/EllipticDiscretizations.f90; EllipticBR1.f90, EllipticBR2.f90, EllipticIP.f90) */
#define H3D_VISCOUS_BR1 0
#define H3D_VISCOUS_BR2 1      /* penaltyParameter = eta (default 2, EllipticBR2.f90:80-91)                          */
#define H3D_VISCOUS_IP 2       /* penaltyParameter = sigma (default 1), ipVariant = SIPG -1 | IIPG 0 | NIPG 1 (EllipticIP.f90:26-28,110-150) */

/* gradient variables (PhysicsStorage_NS.f90:105-108; SpatialDiscretization.f90:106-148; VariableConversion_NS.f90:196-262) */
#define H3D_GRADVARS_STATE 0
#define H3D_GRADVARS_ENTROPY 1
#define H3D_GRADVARS_ENERGY 2

/* face types (libs/mesh/MeshTypes.f90:21-25) */
#define H3D_FACE_INTERIOR 1
#define H3D_FACE_BOUNDARY 2
#define H3D_FACE_MPI 3

/* boundary conditions (libs/physics/common, the BC class files); parameters: 16 doubles per zone */
#define H3D_BC_PERIODIC 0      /* never reaches the device: merged into interior faces at mesh build (HexMesh.f90:519) */
#define H3D_BC_NOSLIPWALL 1    /* [0..2] vWall, [3] wallType (0 adiabatic, 1 isothermal), [4] Twall, [5] T_ref*gammaM2*gammaMinus1, [6] eWall */
#define H3D_BC_FREESLIPWALL 2  /* [3] wallType, [4] Twall, [5] T_ref*gammaM2, [6] eWall                                          */
#define H3D_BC_INFLOW 3        /* [0] rho, [1..3] u,v,w (from |v| and the angles of attack), [4] p; turbulence intensity 0    */
#define H3D_BC_OUTFLOW 4       /* [4] pExt                                                                                   */

/* volume integrals (libs/monitors/VolumeIntegrals.f90:30-60) */
#define H3D_INT_VOLUME 0
#define H3D_INT_KINETIC_ENERGY 1
#define H3D_INT_KINETIC_ENERGY_RATE 2
#define H3D_INT_ENSTROPHY 3
#define H3D_INT_VELOCITY 4          /* "mean velocity",  VolumeIntegrals.f90:288-296 */
#define H3D_INT_ENTROPY 5           /* "entropy",        :298-309                    */
#define H3D_INT_ENTROPY_RATE 6      /* "entropy rate",   :322-332                    */
#define H3D_INT_INTERNAL_ENERGY 7   /* "internal energy",:374-386                    */
#define H3D_INT_ENTROPY_BALANCE 8   /* "entropy balance",:335-370 (entropy rate minus viscous work, per gradient variables) */
#define H3D_INT_MATH_ENTROPY 9      /* "math entropy",   :311-320                    */
#define H3D_INT_KINETIC_ENERGY_BALANCE 10 /* "kinetic energy balance", :220-265 with GetPressureLocalGradient :724-764 (energy gradient variables) */

/* RK schemes (libs/timeintegrator/ExplicitMethods.f90:667,790) */
#define H3D_EULER 1       /* TakeExplicitEulerStep, ExplicitMethods.f90:1232 */
#define H3D_RK3 3
#define H3D_RK5 5
#define H3D_LSERK14_4 14  /* TakeLSERK14_4Step, :884 */
#define H3D_SSPRK33 33     /* TakeSSPRK33Step, :983 (without the optional stage limiter) */
#define H3D_SSPRK43 43     /* TakeSSPRK43Step, :1109 */

/* Run-time physics: the protected module variables the reference's kernels read
 * (libs/physics/navierstokes/PhysicsStorage_NS.f90:81-131,190-250,285,416-429;
 *  FluidData_NS thermodynamics/dimensionless; RiemannSolvers_NS.f90:107-113). */
typedef struct H3dPhysics {
    double gamma;            /* thermodynamics % gamma                          */
    double gammaMinus1;      /* thermodynamics % gammaMinus1  (1.4 - 1.0)       */
    double Mach;             /* dimensionless % Mach                            */
    double Re;               /* dimensionless % Re                              */
    double Pr;               /* dimensionless % Pr                              */
    double mu;               /* dimensionless % mu     = 1/Re                   */
    double kappa;            /* dimensionless % kappa                           */
    double mu_to_kappa;      /* dimensionless % mu_to_kappa                     */
    double gammaM2;          /* dimensionless % gammaM2                         */
    double S_div_Tref;       /* S_div_TRef_Sutherland                           */
    double T_renorm;         /* TemperatureReNormalization_Sutherland           */
    double lambdaStab;       /* RiemannSolvers_NS lambdaStab (0 for central)    */
    double smagorinsky_Cs;   /* "LES model intensity": Smagorinsky CS / WALE Cw / Vreman C */
    double Prt;              /* dimensionless % Prt                             */
    double penaltyParameter; /* BR2 eta / IP sigma ("penalty parameter" key)    */
    int flowIsNavierStokes;  /* 0 = Euler                                       */
    int computeGradients;    /* PhysicsStorage_NS computeGradients              */
    int inviscid;            /* H3D_STANDARD_DG | H3D_SPLIT_DG                  */
    int riemann;             /* H3D_RIEMANN_*                                   */
    int averaging;           /* H3D_AVG_*                                       */
    int les;                 /* H3D_LES_*                                       */
    int les_wall_model;      /* 0 = none (default), 1 = linear: LS = min(Cs*delta, 0.4*dWall), LESModels.f90:189-203 */
    int viscous;             /* H3D_VISCOUS_*                                   */
    int ipVariant;           /* IP: -1 SIPG (default), 0 IIPG, 1 NIPG           */
    int gradientVariables;   /* H3D_GRADVARS_* ("gradient variables" key)       */
} H3dPhysics;

/* ---- life cycle ------------------------------------------------------------------------------
 * replaces MPI_Process % Init / mpi_init (libs/mpiutils/process_info.f90:24-64).  nccl_unique_id is the
 * 128-byte ncclUniqueId broadcast by the host (NULL when nranks == 1). */
int h3d_get_nccl_unique_id(void* id128);
int h3d_create(h3d_handle* out, int rank, int nranks, int device, const void* nccl_unique_id);
int h3d_destroy(h3d_handle h);
const char* h3d_last_error(h3d_handle h);                 /* h may be NULL: creation errors */
int h3d_last_error_copy(h3d_handle h, char* buf, int len); /* Fortran-friendly copy           */

/* ---- set-up (once) ---------------------------------------------------------------------------*/
/* ConstructPhysicsStorage + Initialize_SpaceAndTimeMethods (NavierStokesSolver/main.f90:90,144) */
int h3d_set_physics(h3d_handle h, const H3dPhysics* p);

/* NodalStorage(N) (libs/spectral/NodalStorageClass.f90:24-52).  Matrices row-major M[i*(N+1)+l] = M(i,l);
 * v,b: [side*(N+1)+i], side 0 = LEFT/FRONT/BOTTOM (-1 end), 1 = RIGHT/BACK/TOP (+1 end).  h3d_set_mesh uses the last order given
 * (N = 1..9); every order given stays registered for h3d_set_mesh_p (N = 1..15). */
int h3d_set_basis(h3d_handle h, int N, int nodeType, const double* x, const double* w, const double* D,
                  const double* hatD, const double* sharpD, const double* v, const double* b);

/* HexMesh connectivity + MappedGeometry / MappedGeometryFace (libs/mesh/HexMesh.f90:2399,2677;
 * MappedGeometry.f90:24-55; FaceClass.f90:63-81; HexElementClass.f90:60-75). */
int h3d_set_mesh(h3d_handle h, int nElem, int nFace,
                 const int* elemFace,      /* [nElem][6]  face id of each local face          (e % faceIDs)   */
                 const int* elemFaceSide,  /* [nElem][6]  side of the element on that face     (e % faceSide)  */
                 const int* faceElem,      /* [nFace][2]  element ids, -1 = none               (f % elementIDs)*/
                 const int* faceElemSide,  /* [nFace][2]  local face ids                        (f % elementSide)*/
                 const int* faceRot,       /* [nFace]     0..7                                  (f % rotation)  */
                 const int* faceType,      /* [nFace]     H3D_FACE_*                            (f % faceType)  */
                 const int* faceZone,      /* [nFace]     boundary zone or -1                   (f % zone)      */
                 const double* jGradXi, const double* jGradEta, const double* jGradZeta, /* [e][k][j][i][3] */
                 const double* jacobian,   /* [e][k][j][i]                                                     */
                 const double* x,          /* [e][k][j][i][3]  node coordinates (sources, BCs) - may be NULL   */
                 const double* volume,     /* [e]              e % geom % Volume (LES filter width) - may be NULL */
                 const double* faceNormal, const double* faceT1, const double* faceT2, /* [f][j][i][3]       */
                 const double* faceJacobian, /* [f][j][i]                                                       */
                 const double* faceX,      /* [f][j][i][3]  may be NULL                                        */
                 const double* faceSurface /* [f]           f % geom % surface (LES) - may be NULL             */);

/* ---- p-nonconforming meshes ---------------------------------------------------------------------------------
 * Every element has its own polynomial orders (Nx, Ny, Nz) -- the reference's "polynomial order file" (ReadOrderFile,
 * libs/io/ReadInputFile.f90:132-153) or the result of a p-adaptation -- and every face the orders of its two sides and its own,
 * Nf = max per direction (Face_LinkWithElements, libs/mesh/FaceClass.f90:187-282).  Set-up: h3d_set_physics, then h3d_set_basis once
 * for EVERY order that occurs in an element or on a face (NodalStorage(N), DGSEMClass.f90:215-228; orders up to 15),
 * h3d_set_interpolation for every pair of different orders (N, M) and (M, N) that meet at a face, then h3d_set_mesh_p instead of
 * h3d_set_mesh.  All element arrays of the ABI are then packed element after element at the elements' own sizes, A[e][k][j][i][c]
 * with (Nx+1)(Ny+1)(Nz+1) nodes in element e (StorageClass.f90:423-429), and face arrays face after face at the face orders.
 * Available on such a mesh: StandardDG and SplitDG (every two-point flux; Gauss-Lobatto nodes) with BR1 or the interior penalty
 * (h3d_set_face_h; penalty with maxval(f % Nf), EllipticIP.f90:678-687) (or Euler), every Riemann
 * solver, boundary condition and gradient-variable set, the LES models (filter widths from the element's / face's own orders,
 * SpatialDiscretization.f90:420,1378; h3d_set_wall_distance in the packed sizes),
 * all Runge-Kutta schemes with the stage limiter, h3d_max_residuals / _max_timestep / _has_nan, h3d_volume_integral (volume,
 * kinetic energy and its rate, enstrophy, mean velocity, internal energy, entropy, math entropy, entropy rate, entropy and
 * kinetic-energy balance),
 * h3d_surface_integral, h3d_probe (Lagrange vectors padded to rows of max(N)+1 values), h3d_statistics_*, and h3d_set_halo on
 * partitioned meshes: the traces of the MPI faces are exchanged at the face order.  Refused with a message: BR2.
 * h3d_snapshot_begin takes its copy synchronously on such meshes. */

/* Tset(Norigin, Ndest) % T (libs/spectral/InterpolationMatrices.f90:42-107): row-major T[i*(Norigin+1) + l] = T(i,l),
 * (Ndest+1) x (Norigin+1); Lagrange interpolation for Norigin < Ndest, the L2 projection (weighted transpose) otherwise. */
int h3d_set_interpolation(h3d_handle h, int Norigin, int Ndest, const double* T);

/* h3d_set_mesh for a p-nonconforming mesh: elemOrder[nElem][3] = e % Nxyz (HexElementClass.f90:60-75); the other arguments as
 * h3d_set_mesh with the packed sizes described above (face geometry from MappedGeometryFace at the face order,
 * MappedGeometry.f90:482-756). */
int h3d_set_mesh_p(h3d_handle h, int nElem, int nFace, const int* elemOrder,
                   const int* faceOrder,     /* [nFace][6] f % Nf(1:2), f % NfLeft(1:2), f % NfRight(1:2) (FaceClass.f90:69-71); may be NULL on a
                                                single rank (derived from the elements); required with MPI faces, whose remote element belongs to
                                                another rank (the reference exchanges it: HexMesh_UpdateMPIFacesPolynomial, HexMesh.f90:2425) */
                   const int* elemFace, const int* elemFaceSide, const int* faceElem, const int* faceElemSide,
                   const int* faceRot, const int* faceType, const int* faceZone,
                   const double* jGradXi, const double* jGradEta, const double* jGradZeta, const double* jacobian,
                   const double* x, const double* volume, const double* faceNormal, const double* faceT1, const double* faceT2,
                   const double* faceJacobian, const double* faceX, const double* faceSurface);

/* e % geom % dWall, f % geom % dWall (HexMesh_ComputeWallDistances, libs/mesh/HexMesh.f90:5594-5692): distance to the
 * nearest no-slip wall node, [e][k][j][i] and [f][j][i].  Needed only with les_wall_model = 1. */
int h3d_set_wall_distance(h3d_handle h, const double* dWallElem, const double* dWallFace);

/* f % geom % h (HexMesh.f90:3016-3041, for MPI faces after CommunicateMPIFaceMinimumDistance :3059-3145): the faces'
 * minimum orthogonal distance estimate, [f].  Needed only with viscous = H3D_VISCOUS_IP (penalty, EllipticIP.f90:678-687). */
int h3d_set_face_h(h3d_handle h, const double* faceH);

/* BCs(zone) % bc (libs/physics/common/BoundaryConditions.f90): type + 16 parameters per zone */
int h3d_set_boundary_conditions(h3d_handle h, int nZones, const int* bcType, const double* bcParams);

/* MPIfaces (libs/mpiutils/MPI_Face.f90:16-57; HexMesh.f90:2571-2668).  For each neighbour rank the list of
 * local MPI faces IN EXCHANGE ORDER (both ranks must list the shared faces in the same order) and the side
 * the local element occupies on each. */
int h3d_set_halo(h3d_handle h, int nNeighbors, const int* neighborRank, const int* faceCount,
                 const int* faceIDs, const int* thisSide);

/* ---- state ---------------------------------------------------------------------------------- */
/* global2LocalQ / local2GlobalQ (libs/mesh/StorageClass.f90:390-440,549-579): packed [e][k][j][i][5] */
int h3d_upload_Q(h3d_handle h, const double* Q);
/* any pointer may be NULL.  Ux,Uy,Uz are e % storage % U_x,U_y,U_z of the last residual evaluation. */
int h3d_download(h3d_handle h, double* Q, double* QDot, double* Ux, double* Uy, double* Uz);

/* Autosave without stalling the time loop (TimeIntegrator.f90:924 SaveSolution): _begin takes a
 * device-side snapshot of Q (in the reference's packed layout) at the current point of the time loop and starts its
 * transfer into pinned host memory on a separate stream; the caller keeps stepping; _end waits for the transfer and
 * copies the snapshot into Q.  One snapshot in flight at a time. */
int h3d_snapshot_begin(h3d_handle h);
int h3d_snapshot_end(h3d_handle h, double* Q);
/* e % storage % S_NS evaluated by the host (UserDefinedSourceTermNS, SpatialDiscretization.f90:569-577);
 * NULL clears it.  Kept until replaced. */
int h3d_set_source(h3d_handle h, const double* S);

/* ---- the hot path --------------------------------------------------------------------------- */
/* ComputeTimeDerivative(mesh, particles, time, mode)  (SpatialDiscretization.f90:227-320;
 * interface ComputeTimeDerivative_f, libs/discretization/DGSEMClass.f90:77-95).  Result: QDot (device). */
int h3d_compute_time_derivative(h3d_handle h, double time);

/* TakeRK3Step / TakeRK5Step (libs/timeintegrator/ExplicitMethods.f90:667-788,790-882; interface
 * TimeStep_FCN, TimeIntegratorDefinitions.f90:10-35).  ctd_after_step = CTD_AFTER_STEPS (:784). */
int h3d_rk_step(h3d_handle h, int scheme, double t, double dt, int ctd_after_step);

/* One stage (0-based) of the same schemes: residual at t + b_stage dt with the current source, then
 * G = a_stage G + QDot, Q = Q + c_stage dt G (the loop body of ExplicitMethods.f90:746-760, 857-872).  For
 * time-dependent user source terms (UserDefinedSourceTermNS is evaluated at the stage time,
 * SpatialDiscretization.f90:569-577): the adapter calls h3d_set_source between stages.  h3d_rk_step equals the
 * stages 0..ns-1 in a row. */
int h3d_rk_stage(h3d_handle h, int scheme, int stage, double t, double dt);

/* Enable_limiter + stage_limiter (ExplicitMethods.f90:1737-1847; keys "limit timestep" / "limiter minimum",
 * TimeIntegrator.f90:303-311): after every SSPRK33 / SSPRK43 stage, density and then pressure of each element are scaled
 * towards the element average so that they stay above min(minimum, average).  minimum <= 0 keeps the default 1e-13. */
int h3d_enable_limiter(h3d_handle h, int enabled, double minimum);

/* ---- per-step reductions (all globally reduced over ranks) ----------------------------------- */
/* ComputeMaxResiduals (libs/discretization/DGSEMClass.f90:770-856) */
int h3d_max_residuals(h3d_handle h, double out[5]);
/* MaxTimeStep (DGSEMClass.f90:870-1034): returns both restrictions, caller takes the min */
int h3d_max_timestep(h3d_handle h, double cfl, double dcfl, double* dt_conv, double* dt_visc);
/* ScalarVolumeIntegral (libs/monitors/VolumeIntegrals.f90:76-120,167-286): raw integral, H3D_INT_* */
int h3d_volume_integral(h3d_handle h, int kind, double* val);
/* ScalarSurfaceIntegral / VectorSurfaceIntegral over the faces of boundary zone `zone` (libs/monitors/SurfaceIntegrals.f90:
 * 40-240, 248-445), from the prolonged state of the owning element and, for the viscous kinds, the prolonged gradients of
 * the last residual evaluation (getStressTensor, Physics_NS.f90:822-886).  Raw integrals: the monitors' scalings
 * (SurfaceMonitor.f90:356-446: rho_ref V_ref^2 Lref^2, 2 Lref^2 / reference surface, direction) stay with the caller.
 * out[3]: scalar kinds fill out[0]. */
#define H3D_SURF_SURFACE 0         /* int dS                                   */
#define H3D_SURF_MASS_FLOW 1       /* int rho u.n dS                            */
#define H3D_SURF_FLOW_RATE 2       /* int u.n dS                                */
#define H3D_SURF_PRESSURE 3        /* int p dS (scalar PRESSURE_FORCE: pressure-average numerator) */
#define H3D_SURF_VEC_SURFACE 4     /* int n dS                                  */
#define H3D_SURF_TOTAL_FORCE 5     /* int (p n - tau n) dS                      */
#define H3D_SURF_PRESSURE_FORCE 6  /* int p n dS                                */
#define H3D_SURF_VISCOUS_FORCE 7   /* - int tau n dS                            */
int h3d_surface_integral(h3d_handle h, int zone, int kind, double out[3]);
/* Probe_Update (libs/monitors/Probe.f90:330-420): value = sum_ijk var(i,j,k) lxi(i) leta(j) lzeta(k) in element elem[p] (local
 * 0-based id on this rank), var evaluated at the nodes from Q.  The Lagrange vectors [p][n] come from the host-side
 * point location (HexMesh_FindPointWithCoords, HexMesh.f90:5483; horses3d_b200/probes.py).  Rank-local: no reduction. */
#define H3D_PROBE_PRESSURE 0
#define H3D_PROBE_VELOCITY 1
#define H3D_PROBE_U 2
#define H3D_PROBE_V 3
#define H3D_PROBE_W 4
#define H3D_PROBE_MACH 5
#define H3D_PROBE_K 6
int h3d_probe(h3d_handle h, int nProbes, const int* elem, const int* variable, const double* lxi, const double* leta, const double* lzeta, double* values);
/* StatisticsMonitor_UpdateValues (libs/monitors/StatisticsMonitor.f90:279-540): running averages, kept on the device.
 * Variables per node, in the reference's order: u v w uu vv ww uv uw vw | Q(1:5) | U_x(1:5) U_y(1:5) U_z(1:5) (the gradients
 * only with computeGradients, as "save gradients with solution").  reset != 0 zeroes the averages and the sample count
 * first ("@reset").  download: data[e][k][j][i][nVars] (nVars = 14 or 29), nSamples = samples accumulated so far. */
int h3d_statistics_update(h3d_handle h, int reset);
int h3d_statistics_download(h3d_handle h, double* data, int* nVars, int* nSamples);
/* checkForNan (ExplicitMethods.f90:1856-1905): flag = 1 if any NaN in Q on any rank */
int h3d_has_nan(h3d_handle h, int* flag);

/* ---- utilities ------------------------------------------------------------------------------ */
int h3d_synchronize(h3d_handle h);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long h3d_kernel_launches(h3d_handle h);
/* CUDA-event timing of work enqueued between begin/end on the compute stream, in milliseconds */
int h3d_timer_begin(h3d_handle h);
int h3d_timer_end(h3d_handle h, double* ms);
/* accumulated CUDA-event time per kernel class since the last call (option profile_kernels=1):
 * out = [ms, launches] x {gradient, riemann, volume, prolong} */
int h3d_kernel_profile(h3d_handle h, double* out, int len);
/* Per-stage timeline (option timeline=1): milliseconds since the start of the last residual evaluation of its eleven phase
 * boundaries, on both streams -- 0 start | 1, 2 Q-trace exchange begin / end (communication stream) | 3 gradient(interior
 * elements) end | 4 gradient(MPI elements) end | 5, 6 gradient-trace exchange begin / end (communication stream) | 7 riemann(local
 * faces) end | 8 volume(interior elements) end | 9 riemann(MPI faces) end | 10 volume(MPI elements) end; -1 where a phase did
 * not run.  Replaces the reference's Stopwatch around UpdateMPIFaces* / GatherMPIFaces* (HexMesh.f90:1199-1391). */
int h3d_stage_timeline(h3d_handle h, double* marks_ms, int len);
/* option string "key=value"; unknown keys are an error.  store_qdot_every_stage=1 | profile_kernels=1 | timeline=1 |
 * use_tma=0 (plain-load kernels) | mma=1 (n = 8: contractions on the FP64 tensor cores, not bit-identical to the CUDA-core
 * summation order) | gen2=1 (n = 8: 256-thread kernels, two CTAs per SM) | comm_sms=K, interior_split_pct=P (on a rank with
 * neighbours the first P % of the interior elements run on all but K multiprocessors, which the halo exchange uses meanwhile;
 * defaults 8 and 30) | sync_mpi_face_geometry=0 (keep the geometry each rank built for its MPI faces from its own element, as the
 * reference does; by default the rank that owns the left side of an MPI face sends normal, tangents, surface Jacobian and LES
 * width to its neighbour once, so that a partitioned run uses the face geometry of the single-domain run) */
int h3d_set_option(h3d_handle h, const char* key_value);

#ifdef __cplusplus
}
#endif
#endif /* H3D_GPU_H */
