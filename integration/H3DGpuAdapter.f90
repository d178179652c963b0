!
!   H3DGpuAdapter -- the module a maintainer of HORSES3D adds to Solver/src/libs/timeintegrator/ to run the explicit
!   Navier-Stokes path on libh3dgpu.so (include/h3d_gpu.h).  It is written against integration/h3d_gpu_interfaces.f90
!   (module H3DGpuInterfaces, generated from the header).  Every routine names the reference code it replaces.
!
!   Integration mode: RESIDENT STATE.  The solution lives on the device for the whole time loop; the host copy
!   (mesh % storage % Q and the element storages) is refreshed only where the reference needs it on the host:
!   before SaveSolution (TimeIntegrator.f90:924, main.f90:207) and before host-side monitors that are not offloaded.
!   Per step the driver calls h3d_rk_step, h3d_max_residuals, h3d_volume_integral / h3d_surface_integral for the
!   monitors and h3d_has_nan: a few hundred bytes cross PCIe per step (bench.py "resident": 3.9 G DOF-updates/s
!   against 1.4 G when the state is uploaded and downloaded every step).
!
!   No Fortran compiler exists in the build container of this repository, so this file has been checked by
!   tests/test_cabi_exports.py only for consistency with the generated interface block (every h3d_* name it calls
!   is declared there with the same number of arguments); it has not been compiled.
!
#include "Includes.h"
module H3DGpuAdapter
   use, intrinsic :: iso_c_binding
   use SMConstants
   use HexMeshClass
   use ElementClass
   use FaceClass
   use NodalStorageClass
   use InterpolationMatrices,      only: Tset
   use PhysicsStorage
   use FluidData,                  only: thermodynamics, dimensionless, refValues
   use RiemannSolvers_NS,          only: whichRiemannSolver, whichAverage, lambdaStab, &
                                         RIEMANN_ROE, RIEMANN_LXF, RIEMANN_RUSANOV, RIEMANN_STDROE, RIEMANN_CENTRAL, &
                                         RIEMANN_ROEPIKE, RIEMANN_LOWDISSROE, RIEMANN_MATRIXDISS, RIEMANN_UDISS, &
                                         STANDARD_AVG, MORINISHI_AVG, DUCROS_AVG, KENNEDYGRUBER_AVG, PIROZZOLI_AVG, &
                                         ENTROPYCONS_AVG, CHANDRASEKAR_AVG
   use LESModels,                  only: LESModel, Smagorinsky_t, WALE_t, Vreman_t
   use BoundaryConditions,         only: BCs
   use InflowBCClass,              only: InflowBC_t
   use OutflowBCClass,             only: OutflowBC_t
   use NoSlipWallBCClass,          only: NoSlipWallBC_t
   use FreeSlipWallBCClass,        only: FreeSlipWallBC_t
   use MPI_Process_Info,           only: MPI_Process
   use ParticlesClass,             only: Particles_t
   use H3DGpuInterfaces
#ifdef _HAS_MPI_
   use mpi
#endif
   implicit none
   private
   public  h3d_gpu_initialize, h3d_gpu_setup, h3d_gpu_finalize, h3d_gpu_download_state
   public  ComputeTimeDerivative_GPU, TakeExplicitEulerStep_GPU, TakeRK3Step_GPU, TakeRK5Step_GPU
   public  TakeLSERK14_4Step_GPU, TakeSSPRK33Step_GPU, TakeSSPRK43Step_GPU
   public  ComputeMaxResiduals_GPU, MaxTimeStep_GPU, ScalarVolumeIntegral_GPU, VectorSurfaceIntegral_GPU, checkForNan_GPU

   type(c_ptr), save :: h3d = c_null_ptr
   logical,     save :: ctd_after_steps = .false.       ! CTD_AFTER_STEPS of ExplicitMethods.f90:32

contains
!
!  ----------------------------------------------------------------------------------------------------------------
!  Error convention of the reference: errorMessage(STD_OUT) + error stop (Includes.h:5)
!  ----------------------------------------------------------------------------------------------------------------
   subroutine check(rc, where)
      integer(c_int),   intent(in) :: rc
      character(len=*), intent(in) :: where
      character(kind=c_char)       :: buf(512)
      integer                      :: k, l
      if ( rc == 0 ) return
      l = h3d_last_error_copy(h3d, buf, 512_c_int)
      write(STD_OUT,'(A,A,A,I0,A)',advance="no") "h3d: ", where, " failed (", rc, "): "
      do k = 1, 512
         if ( buf(k) == c_null_char ) exit
         write(STD_OUT,'(A1)',advance="no") buf(k)
      end do
      write(STD_OUT,*)
      errorMessage(STD_OUT)
      error stop
   end subroutine check
!
!  ----------------------------------------------------------------------------------------------------------------
!  One rank <-> one GPU.  Called right after MPI_Process % Init (main.f90:60): rank 0 makes the NCCL id, everybody
!  gets it by MPI_Bcast, every rank opens its context.
!  ----------------------------------------------------------------------------------------------------------------
   subroutine h3d_gpu_initialize(gpus_per_node)
      integer, intent(in)            :: gpus_per_node
      character(kind=c_char), target :: id(128)
      integer                        :: ierr
      id = c_null_char
      if ( MPI_Process % isRoot ) call check(h3d_get_nccl_unique_id(c_loc(id)), "h3d_get_nccl_unique_id")
#ifdef _HAS_MPI_
      if ( MPI_Process % doMPIAction ) call mpi_bcast(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
#endif
      call check(h3d_create(h3d, int(MPI_Process % rank, c_int), int(MPI_Process % nProcs, c_int), &
                            int(mod(MPI_Process % rank, gpus_per_node), c_int), c_loc(id)), "h3d_create")
   end subroutine h3d_gpu_initialize

   subroutine h3d_gpu_finalize()
      if ( c_associated(h3d) ) call check(h3d_destroy(h3d), "h3d_destroy")
      h3d = c_null_ptr
   end subroutine h3d_gpu_finalize
!
!  ----------------------------------------------------------------------------------------------------------------
!  Once, after sem % construct and Initialize_SpaceAndTimeMethods (main.f90:117-144): physics, basis, mesh, boundary
!  table, wall distances, halo, state.  A mesh whose elements do not all have the order N (a polynomial order file, a
!  p-adapted mesh) takes the p-nonconforming branch (setup_basis_and_mesh_p, below).
!  ----------------------------------------------------------------------------------------------------------------
   subroutine h3d_gpu_setup(mesh, N, viscousDiscretization, penaltyParameter, ipVariant, gradientVariables, inviscidIsSplitForm, ctdAfterSteps)
      type(HexMesh), target, intent(inout) :: mesh
      integer,               intent(in)    :: N
      integer,               intent(in)    :: viscousDiscretization      ! H3D_VISCOUS_*  (SpatialDiscretization.f90:106-195)
      real(kind=RP),         intent(in)    :: penaltyParameter
      integer,               intent(in)    :: ipVariant, gradientVariables
      logical,               intent(in)    :: inviscidIsSplitForm, ctdAfterSteps
      type(H3dPhysics)                     :: p
      integer                              :: eID, fID, s, k, zID, nE, nF, n3, n2, dom, nShared, pos, nZones
      integer(c_int), allocatable          :: elemFace(:), elemFaceSide(:), faceElem(:), faceElemSide(:), faceRot(:), faceType(:), faceZone(:)
      real(c_double), allocatable          :: jGradXi(:), jGradEta(:), jGradZeta(:), jac(:), x(:), vol(:)
      real(c_double), allocatable          :: fN(:), fT1(:), fT2(:), fJ(:), fX(:), fS(:), fH(:), dWe(:), dWf(:)
      integer(c_int), allocatable          :: bcType(:), nbrRank(:), nbrCount(:), haloFace(:), haloSide(:)
      real(c_double), allocatable          :: bcPar(:), Dt(:), hatDt(:), sharpDt(:), vv(:), bb(:)
      type(NodalStorage_t), pointer        :: sp
      logical                              :: isMixed
      integer                              :: ierr

      ctd_after_steps = ctdAfterSteps
!
!     1. Physics: thermodynamics, dimensionless numbers, Sutherland constants, solver choices (PhysicsStorage_NS.f90:190-306)
!     ------------------------------------------------------------------------------------------------------------------
      p % gamma = thermodynamics % gamma;        p % gammaMinus1 = thermodynamics % gammaMinus1
      p % Mach  = dimensionless % Mach;          p % Re = dimensionless % Re;   p % Pr = dimensionless % Pr
      p % mu    = dimensionless % mu;            p % kappa = dimensionless % kappa
      p % mu_to_kappa = dimensionless % mu_to_kappa;   p % gammaM2 = dimensionless % gammaM2;   p % Prt = dimensionless % Prt
      p % S_div_Tref = S_div_TRef_Sutherland;    p % T_renorm = TemperatureReNormalization_Sutherland
      p % lambdaStab = lambdaStab
      p % penaltyParameter = penaltyParameter
      p % flowIsNavierStokes = merge(1_c_int, 0_c_int, flowIsNavierStokes)
      p % computeGradients   = merge(1_c_int, 0_c_int, computeGradients)
      p % inviscid = merge(H3D_SPLIT_DG, H3D_STANDARD_DG, inviscidIsSplitForm)
      select case ( whichRiemannSolver )
      case (RIEMANN_ROE);        p % riemann = H3D_RIEMANN_ROE
      case (RIEMANN_LXF);        p % riemann = H3D_RIEMANN_LXF
      case (RIEMANN_CENTRAL);    p % riemann = H3D_RIEMANN_CENTRAL
      case (RIEMANN_RUSANOV);    p % riemann = H3D_RIEMANN_RUSANOV
      case (RIEMANN_STDROE);     p % riemann = H3D_RIEMANN_STDROE
      case (RIEMANN_UDISS);      p % riemann = H3D_RIEMANN_UDISS
      case (RIEMANN_ROEPIKE);    p % riemann = H3D_RIEMANN_ROEPIKE
      case (RIEMANN_LOWDISSROE); p % riemann = H3D_RIEMANN_LOWDISSROE
      case (RIEMANN_MATRIXDISS); p % riemann = H3D_RIEMANN_MATRIXDISS
      case default
         print*, "Riemann Solver not recognized by the GPU path."
         errorMessage(STD_OUT) ; error stop
      end select
      select case ( whichAverage )
      case (STANDARD_AVG);      p % averaging = H3D_AVG_STANDARD
      case (KENNEDYGRUBER_AVG); p % averaging = H3D_AVG_KENNEDYGRUBER
      case (PIROZZOLI_AVG);     p % averaging = H3D_AVG_PIROZZOLI
      case (DUCROS_AVG);        p % averaging = H3D_AVG_DUCROS
      case (MORINISHI_AVG);     p % averaging = H3D_AVG_MORINISHI
      case (ENTROPYCONS_AVG);   p % averaging = H3D_AVG_ENTROPYCONS
      case (CHANDRASEKAR_AVG);  p % averaging = H3D_AVG_CHANDRASEKAR
      case default;             p % averaging = H3D_AVG_STANDARD
      end select
      p % les = H3D_LES_NONE;  p % smagorinsky_Cs = 0.0_RP;  p % les_wall_model = 0_c_int
      if ( allocated(LESModel) ) then
         select type ( m => LESModel )                                 ! LESModels.f90:49-71
         type is (Smagorinsky_t); p % les = H3D_LES_SMAGORINSKY;  p % smagorinsky_Cs = m % CS
         type is (WALE_t);        p % les = H3D_LES_WALE;         p % smagorinsky_Cs = m % Cw
         type is (Vreman_t);      p % les = H3D_LES_VREMAN;       p % smagorinsky_Cs = m % C
         end select
         if ( LESModel % WallModel == 1 ) p % les_wall_model = 1_c_int   ! LINEAR_WALLMODEL, LESModels.f90:137-165
      end if
      p % viscous = int(viscousDiscretization, c_int);  p % ipVariant = int(ipVariant, c_int)
      p % gradientVariables = int(gradientVariables, c_int)
      call check(h3d_set_physics(h3d, p), "h3d_set_physics")
!
!     2. Basis: the header wants row-major M(i,l) -> the transposes of the Fortran column-major arrays
!     ----------------------------------------------------------------------------------------------
      isMixed = any(mesh % Nx /= N) .or. any(mesh % Ny /= N) .or. any(mesh % Nz /= N)
#ifdef _HAS_MPI_
      if ( MPI_Process % doMPIAction ) call mpi_allreduce(MPI_IN_PLACE, isMixed, 1, MPI_LOGICAL, MPI_LOR, MPI_COMM_WORLD, ierr)   ! one decision for all ranks
#endif
      if ( isMixed ) then
         call setup_basis_and_mesh_p(mesh)
         goto 400                                                          ! boundary table, state
      end if
      sp => NodalStorage(N)
      allocate(Dt((N+1)**2), hatDt((N+1)**2), sharpDt((N+1)**2), vv(2*(N+1)), bb(2*(N+1)))
      Dt      = reshape(transpose(sp % D),      [(N+1)**2])
      hatDt   = reshape(transpose(sp % hatD),   [(N+1)**2])
      sharpDt = 0.0_RP
      if ( allocated(sp % sharpD) ) sharpDt = reshape(transpose(sp % sharpD), [(N+1)**2])
      vv = reshape(sp % v, [2*(N+1)]);   bb = reshape(sp % b, [2*(N+1)])            ! v(0:N,side): side-major as the header asks
      call check(h3d_set_basis(h3d, int(N, c_int), int(mesh % nodeType, c_int), sp % x, sp % w, Dt, hatDt, sharpDt, vv, bb), "h3d_set_basis")
!
!     3. Mesh: connectivity (0-based) and geometry, elements and faces in local ID order
!     -----------------------------------------------------------------------------------
      nE = size(mesh % elements);  nF = size(mesh % faces);  n3 = (N+1)**3;  n2 = (N+1)**2
      allocate(elemFace(6*nE), elemFaceSide(6*nE), faceElem(2*nF), faceElemSide(2*nF), faceRot(nF), faceType(nF), faceZone(nF))
      allocate(jGradXi(3*n3*nE), jGradEta(3*n3*nE), jGradZeta(3*n3*nE), jac(n3*nE), x(3*n3*nE), vol(nE))
      allocate(fN(3*n2*nF), fT1(3*n2*nF), fT2(3*n2*nF), fJ(n2*nF), fX(3*n2*nF), fS(nF), fH(nF))
      do eID = 1, nE
         associate ( e => mesh % elements(eID) )
         if ( any(e % Nxyz /= N) ) then
            print*, "The GPU path needs one polynomial order for the whole mesh."
            errorMessage(STD_OUT) ; error stop
         end if
         do s = 1, 6
            elemFace(6*(eID-1)+s)     = e % faceIDs(s) - 1                     ! HexElementClass.f90:67
            elemFaceSide(6*(eID-1)+s) = e % faceSide(s) - 1                    ! 0 = left, 1 = right
         end do
         jGradXi  (3*n3*(eID-1)+1 : 3*n3*eID) = reshape(e % geom % jGradXi,   [3*n3])     ! (3,0:N,0:N,0:N): component fastest
         jGradEta (3*n3*(eID-1)+1 : 3*n3*eID) = reshape(e % geom % jGradEta,  [3*n3])
         jGradZeta(3*n3*(eID-1)+1 : 3*n3*eID) = reshape(e % geom % jGradZeta, [3*n3])
         x        (3*n3*(eID-1)+1 : 3*n3*eID) = reshape(e % geom % x,         [3*n3])
         jac      (  n3*(eID-1)+1 :   n3*eID) = reshape(e % geom % jacobian,  [n3])
         vol(eID) = e % geom % volume
         end associate
      end do
      do fID = 1, nF
         associate ( f => mesh % faces(fID) )
         do k = 1, 2
            faceElem(2*(fID-1)+k)     = f % elementIDs(k) - 1                  ! HMESH_NONE = 0 -> -1
            faceElemSide(2*(fID-1)+k) = f % elementSide(k) - 1
         end do
         faceRot(fID) = f % rotation;  faceType(fID) = f % faceType;  faceZone(fID) = f % zone - 1
         fN (3*n2*(fID-1)+1 : 3*n2*fID) = reshape(f % geom % normal,   [3*n2])
         fT1(3*n2*(fID-1)+1 : 3*n2*fID) = reshape(f % geom % t1,       [3*n2])
         fT2(3*n2*(fID-1)+1 : 3*n2*fID) = reshape(f % geom % t2,       [3*n2])
         fX (3*n2*(fID-1)+1 : 3*n2*fID) = reshape(f % geom % x,        [3*n2])
         fJ (  n2*(fID-1)+1 :   n2*fID) = reshape(f % geom % jacobian, [n2])
         fS(fID) = f % geom % surface;  fH(fID) = f % geom % h
         end associate
      end do
      call check(h3d_set_mesh(h3d, int(nE, c_int), int(nF, c_int), elemFace, elemFaceSide, faceElem, faceElemSide, faceRot, faceType, faceZone, &
                              jGradXi, jGradEta, jGradZeta, jac, x, vol, fN, fT1, fT2, fJ, fX, fS), "h3d_set_mesh")
!
!     4. Boundary table: one type and 16 parameters per zone (include/h3d_gpu.h, H3D_BC_*)
!     --------------------------------------------------------------------------------------
400   continue
      if ( viscousDiscretization == H3D_VISCOUS_IP ) then                ! f % geom % h (HexMesh.f90:3016-3041), uniform or not
         if ( .not. allocated(fH) ) allocate(fH(size(mesh % faces)))
         do fID = 1, size(mesh % faces) ; fH(fID) = mesh % faces(fID) % geom % h ; end do
         call check(h3d_set_face_h(h3d, fH), "h3d_set_face_h")
      end if
      nZones = size(mesh % zones)
      if ( nZones > 0 ) then
         allocate(bcType(nZones), bcPar(16*nZones));  bcPar = 0.0_RP
         do zID = 1, nZones
            pos = 16*(zID-1)
            select type ( bc => BCs(zID) % bc )
            type is (NoSlipWallBC_t)                                          ! NoSlipWallBC.f90:150-210
               bcType(zID) = H3D_BC_NOSLIPWALL
               bcPar(pos+1:pos+3) = bc % vWall;  bcPar(pos+4) = bc % wallType;  bcPar(pos+5) = bc % Twall
               bcPar(pos+6) = refValues % T * dimensionless % gammaM2 * thermodynamics % gammaMinus1;  bcPar(pos+7) = bc % ewall
            type is (FreeSlipWallBC_t)                                        ! FreeSlipWallBC.f90:150-200
               bcType(zID) = H3D_BC_FREESLIPWALL
               bcPar(pos+4) = bc % wallType;  bcPar(pos+5) = bc % Twall
               bcPar(pos+6) = refValues % T * dimensionless % gammaM2;  bcPar(pos+7) = bc % ewall
            type is (InflowBC_t)                                              ! InflowBC.f90:380-401, turbulence intensity 0
               bcType(zID) = H3D_BC_INFLOW
               bcPar(pos+1) = bc % rho
               bcPar(pos+2) = bc % v * cos(bc % AoATheta) * cos(bc % AoAPhi)
               bcPar(pos+3) = bc % v * sin(bc % AoATheta) * cos(bc % AoAPhi)
               bcPar(pos+4) = bc % v * sin(bc % AoAPhi)
               bcPar(pos+5) = bc % p
               if ( bc % TurbIntensity /= 0.0_RP ) then
                  print*, "The GPU path has no synthetic inflow turbulence." ; errorMessage(STD_OUT) ; error stop
               end if
            type is (OutflowBC_t)                                             ! OutflowBC.f90:226-288
               bcType(zID) = H3D_BC_OUTFLOW;  bcPar(pos+5) = bc % pExt
            class default                                                     ! periodic zones were merged into interior faces
               bcType(zID) = H3D_BC_PERIODIC
            end select
         end do
         call check(h3d_set_boundary_conditions(h3d, int(nZones, c_int), bcType, bcPar), "h3d_set_boundary_conditions")
      end if
!
!     LES wall model: distances computed by the reference over ALL ranks (HexMesh.f90:5594-5780)
      if ( p % les_wall_model == 1 ) then                               ! packed at the elements' / faces' own sizes (uniform or not)
         nE = size(mesh % elements);  nF = size(mesh % faces)
         pos = 0
         do eID = 1, nE ; pos = pos + size(mesh % elements(eID) % geom % dWall) ; end do
         allocate(dWe(pos))
         pos = 0
         do fID = 1, nF ; pos = pos + size(mesh % faces(fID) % geom % dWall) ; end do
         allocate(dWf(pos))
         pos = 0
         do eID = 1, nE
            k = size(mesh % elements(eID) % geom % dWall)
            dWe(pos+1 : pos+k) = reshape(mesh % elements(eID) % geom % dWall, [k]);  pos = pos + k
         end do
         pos = 0
         do fID = 1, nF
            k = size(mesh % faces(fID) % geom % dWall)
            dWf(pos+1 : pos+k) = reshape(mesh % faces(fID) % geom % dWall, [k]);  pos = pos + k
         end do
         call check(h3d_set_wall_distance(h3d, dWe, dWf), "h3d_set_wall_distance")
      end if
!
!     5. Halo: the order of MPIfaces % faces(domain) % faceIDs is what makes send and receive orders match (HexMesh.f90:2614-2664)
!     ------------------------------------------------------------------------------------------------------------------------
      if ( MPI_Process % doMPIAction ) then
         nShared = 0
         do dom = 1, MPI_Process % nProcs
            if ( mesh % MPIfaces % faces(dom) % no_of_faces > 0 ) nShared = nShared + 1
         end do
         allocate(nbrRank(nShared), nbrCount(nShared), haloFace(sum(mesh % MPIfaces % faces(:) % no_of_faces)), haloSide(size(haloFace)))
         k = 0;  pos = 0
         do dom = 1, MPI_Process % nProcs
            associate ( mf => mesh % MPIfaces % faces(dom) )
            if ( mf % no_of_faces == 0 ) cycle
            k = k + 1;  nbrRank(k) = dom - 1;  nbrCount(k) = mf % no_of_faces
            haloFace(pos+1 : pos+mf % no_of_faces) = mf % faceIDs(1:mf % no_of_faces) - 1
            haloSide(pos+1 : pos+mf % no_of_faces) = mf % elementSide(1:mf % no_of_faces) - 1
            pos = pos + mf % no_of_faces
            end associate
         end do
         call check(h3d_set_halo(h3d, int(nShared, c_int), nbrRank, nbrCount, haloFace, haloSide), "h3d_set_halo")
      end if
!
!     6. State: the reference's packed global vector (StorageClass.f90:423-429)
!     ----------------------------------------------------------------------------
      call mesh % storage % local2GlobalQ(mesh % storage % NDOF)
      call check(h3d_upload_Q(h3d, mesh % storage % Q), "h3d_upload_Q")
   end subroutine h3d_gpu_setup
!
!  ----------------------------------------------------------------------------------------------------------------
!  Steps 2 and 3 of the set-up on a p-nonconforming mesh (include/h3d_gpu.h, h3d_set_mesh_p): NodalStorage(N) of every
!  constructed order (DGSEMClass.f90:215-228, FaceClass.f90:226-232), Tset(N,M) of every constructed pair of orders
!  (FaceClass.f90:236-251), and the mesh with the elements' orders; arrays packed at the elements' / faces' own sizes.
!  ----------------------------------------------------------------------------------------------------------------
   subroutine setup_basis_and_mesh_p(mesh)
      type(HexMesh), target, intent(inout) :: mesh
      integer                              :: N, M, eID, fID, s, k, nE, nF, n3, n2, posE, posF
      integer(c_int), allocatable          :: elemOrder(:), faceOrder(:), elemFace(:), elemFaceSide(:), faceElem(:), faceElemSide(:), faceRot(:), faceType(:), faceZone(:)
      real(c_double), allocatable          :: jGradXi(:), jGradEta(:), jGradZeta(:), jac(:), x(:), vol(:)
      real(c_double), allocatable          :: fN(:), fT1(:), fT2(:), fJ(:), fX(:), fS(:)
      real(c_double), allocatable          :: Dt(:), hatDt(:), sharpDt(:), vv(:), bb(:), Tt(:)
      type(NodalStorage_t), pointer        :: sp

      do N = 1, ubound(NodalStorage, 1)
         if ( .not. NodalStorage(N) % Constructed ) cycle
         sp => NodalStorage(N)
         allocate(Dt((N+1)**2), hatDt((N+1)**2), sharpDt((N+1)**2), vv(2*(N+1)), bb(2*(N+1)))
         Dt      = reshape(transpose(sp % D),    [(N+1)**2])
         hatDt   = reshape(transpose(sp % hatD), [(N+1)**2])
         sharpDt = 0.0_RP
         vv = reshape(sp % v, [2*(N+1)]);   bb = reshape(sp % b, [2*(N+1)])
         call check(h3d_set_basis(h3d, int(N, c_int), int(mesh % nodeType, c_int), sp % x, sp % w, Dt, hatDt, sharpDt, vv, bb), "h3d_set_basis")
         deallocate(Dt, hatDt, sharpDt, vv, bb)
      end do
      do M = 1, ubound(Tset, 2) ; do N = 1, ubound(Tset, 1)                  ! Tset(Norigin, Ndest) % T(0:Ndest, 0:Norigin)
         if ( N == M ) cycle
         if ( .not. Tset(N, M) % Constructed ) cycle
         allocate(Tt((N+1)*(M+1)))
         Tt = reshape(transpose(Tset(N, M) % T), [(N+1)*(M+1)])             ! row-major T(i,l) as the header asks
         call check(h3d_set_interpolation(h3d, int(N, c_int), int(M, c_int), Tt), "h3d_set_interpolation")
         deallocate(Tt)
      end do                    ; end do

      nE = size(mesh % elements);  nF = size(mesh % faces)
      posE = 0
      do eID = 1, nE ; posE = posE + product(mesh % elements(eID) % Nxyz + 1) ; end do
      posF = 0
      do fID = 1, nF ; posF = posF + product(mesh % faces(fID) % Nf + 1) ; end do
      allocate(elemOrder(3*nE), faceOrder(6*nF), elemFace(6*nE), elemFaceSide(6*nE), faceElem(2*nF), faceElemSide(2*nF), faceRot(nF), faceType(nF), faceZone(nF))
      allocate(jGradXi(3*posE), jGradEta(3*posE), jGradZeta(3*posE), jac(posE), x(3*posE), vol(nE))
      allocate(fN(3*posF), fT1(3*posF), fT2(3*posF), fJ(posF), fX(3*posF), fS(nF))
      posE = 0
      do eID = 1, nE
         associate ( e => mesh % elements(eID) )
         n3 = product(e % Nxyz + 1)
         elemOrder(3*(eID-1)+1 : 3*eID) = e % Nxyz
         do s = 1, 6
            elemFace(6*(eID-1)+s)     = e % faceIDs(s) - 1
            elemFaceSide(6*(eID-1)+s) = e % faceSide(s) - 1
         end do
         jGradXi  (3*posE+1 : 3*(posE+n3)) = reshape(e % geom % jGradXi,   [3*n3])
         jGradEta (3*posE+1 : 3*(posE+n3)) = reshape(e % geom % jGradEta,  [3*n3])
         jGradZeta(3*posE+1 : 3*(posE+n3)) = reshape(e % geom % jGradZeta, [3*n3])
         x        (3*posE+1 : 3*(posE+n3)) = reshape(e % geom % x,         [3*n3])
         jac      (  posE+1 :    posE+n3 ) = reshape(e % geom % jacobian,  [n3])
         vol(eID) = e % geom % volume
         posE = posE + n3
         end associate
      end do
      posF = 0
      do fID = 1, nF
         associate ( f => mesh % faces(fID) )
         n2 = product(f % Nf + 1)
         do k = 1, 2
            faceElem(2*(fID-1)+k)     = f % elementIDs(k) - 1
            faceElemSide(2*(fID-1)+k) = f % elementSide(k) - 1
         end do
         faceRot(fID) = f % rotation;  faceType(fID) = f % faceType;  faceZone(fID) = f % zone - 1
         faceOrder(6*(fID-1)+1 : 6*fID) = [f % Nf, f % NfLeft, f % NfRight]       ! MPI faces: after UpdateMPIFacesPolynomial (HexMesh.f90:2425)
         fN (3*posF+1 : 3*(posF+n2)) = reshape(f % geom % normal,   [3*n2])
         fT1(3*posF+1 : 3*(posF+n2)) = reshape(f % geom % t1,       [3*n2])
         fT2(3*posF+1 : 3*(posF+n2)) = reshape(f % geom % t2,       [3*n2])
         fX (3*posF+1 : 3*(posF+n2)) = reshape(f % geom % x,        [3*n2])
         fJ (  posF+1 :    posF+n2 ) = reshape(f % geom % jacobian, [n2])
         fS(fID) = f % geom % surface
         posF = posF + n2
         end associate
      end do
      call check(h3d_set_mesh_p(h3d, int(nE, c_int), int(nF, c_int), elemOrder, faceOrder, elemFace, elemFaceSide, faceElem, faceElemSide, faceRot, faceType, &
                                faceZone, jGradXi, jGradEta, jGradZeta, jac, x, vol, fN, fT1, fT2, fJ, fX, fS), "h3d_set_mesh_p")
   end subroutine setup_basis_and_mesh_p
!
!  ----------------------------------------------------------------------------------------------------------------
!  Host copy of the state for SaveSolution / restart / host-side monitors (TimeIntegrator.f90:924, main.f90:207)
!  ----------------------------------------------------------------------------------------------------------------
   subroutine h3d_gpu_download_state(mesh, withQDot)
      type(HexMesh), target, intent(inout) :: mesh
      logical,               intent(in)    :: withQDot
      if ( withQDot ) then
         call check(h3d_download(h3d, c_loc(mesh % storage % Q), c_loc(mesh % storage % QDot), c_null_ptr, c_null_ptr, c_null_ptr), "h3d_download")
         call mesh % storage % global2LocalQdot
      else
         call check(h3d_download(h3d, c_loc(mesh % storage % Q), c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr), "h3d_download")
      end if
      call mesh % storage % global2LocalQ
   end subroutine h3d_gpu_download_state
!
!  ----------------------------------------------------------------------------------------------------------------
!  ComputeTimeDerivative_f (DGSEMClass.f90:77-95), passed to timeIntegrator % integrate at main.f90:169
!  ----------------------------------------------------------------------------------------------------------------
   subroutine ComputeTimeDerivative_GPU( mesh, particles, time, mode, HO_Elements, element_mask, Level )
      type(HexMesh), target           :: mesh
      type(Particles_t)               :: particles
      real(kind=RP)                   :: time
      integer,             intent(in) :: mode
      logical, intent(in), optional   :: HO_Elements
      logical, intent(in), optional   :: element_mask(:)
      integer, intent(in), optional   :: Level
      if ( present(element_mask) .or. present(Level) ) then
         print*, "The GPU residual evaluates every element: element masks / multi-level RK are not offloaded."
         errorMessage(STD_OUT) ; error stop
      end if
      call check(h3d_compute_time_derivative(h3d, time), "h3d_compute_time_derivative")
   end subroutine ComputeTimeDerivative_GPU
!
!  ----------------------------------------------------------------------------------------------------------------
!  TimeStep_FCN (TimeIntegratorDefinitions.f90:10-35): self % RKStep => Take...Step_GPU at TimeIntegrator.f90:229-249
!  ----------------------------------------------------------------------------------------------------------------
   subroutine take_step(scheme, t, deltaT, dt_vec, dts, global_dt, dtAdaptation)
      integer(c_int),          intent(in) :: scheme
      real(kind=RP),           intent(in) :: t, deltaT
      real(kind=RP), optional, intent(in) :: dt_vec(:)
      logical,       optional, intent(in) :: dts, dtAdaptation
      real(kind=RP), optional, intent(in) :: global_dt
      if ( present(dt_vec) ) then
         print*, "Local time stepping is not offloaded to the GPU." ; errorMessage(STD_OUT) ; error stop
      end if
      if ( present(dts) ) then
         if ( dts ) then
            print*, "Dual time stepping is not offloaded to the GPU." ; errorMessage(STD_OUT) ; error stop
         end if
      end if
      call check(h3d_rk_step(h3d, scheme, t, deltaT, merge(1_c_int, 0_c_int, ctd_after_steps)), "h3d_rk_step")
   end subroutine take_step

#define H3D_STEPPER(NAME, SCHEME) \
   subroutine NAME( mesh, particles, t, deltaT, ComputeTimeDerivative, dt_vec, dts, global_dt, iter, dtAdaptation ) ; \
      type(HexMesh) :: mesh ; type(Particles_t) :: particles ; real(kind=RP) :: t, deltaT ; \
      procedure(ComputeTimeDerivative_f) :: ComputeTimeDerivative ; \
      real(kind=RP), allocatable, dimension(:), intent(in), optional :: dt_vec ; \
      logical, intent(in), optional :: dts ; real(kind=RP), intent(in), optional :: global_dt ; \
      integer, intent(in), optional :: iter ; logical, intent(in), optional :: dtAdaptation ; \
      call take_step(SCHEME, t, deltaT, dt_vec, dts, global_dt, dtAdaptation) ; \
   end subroutine NAME

   H3D_STEPPER(TakeExplicitEulerStep_GPU, H3D_EULER)
   H3D_STEPPER(TakeRK3Step_GPU,           H3D_RK3)
   H3D_STEPPER(TakeRK5Step_GPU,           H3D_RK5)
   H3D_STEPPER(TakeLSERK14_4Step_GPU,     H3D_LSERK14_4)
   H3D_STEPPER(TakeSSPRK33Step_GPU,       H3D_SSPRK33)
   H3D_STEPPER(TakeSSPRK43Step_GPU,       H3D_SSPRK43)
!
!  ----------------------------------------------------------------------------------------------------------------
!  The per-step reductions (all already reduced over the ranks by the library)
!  ----------------------------------------------------------------------------------------------------------------
   function ComputeMaxResiduals_GPU() result(maxResidual)                   ! DGSEMClass.f90:770-856
      real(kind=RP) :: maxResidual(NCONS)
      call check(h3d_max_residuals(h3d, maxResidual), "h3d_max_residuals")
   end function ComputeMaxResiduals_GPU

   subroutine MaxTimeStep_GPU(cfl, dcfl, MaxDt, MaxDtVec)                   ! DGSEMClass.f90:870-1034
      real(kind=RP),           intent(in)  :: cfl, dcfl
      real(kind=RP),           intent(out) :: MaxDt
      real(kind=RP), optional, intent(out) :: MaxDtVec(:)
      real(kind=RP) :: dt_conv, dt_visc
      if ( present(MaxDtVec) ) then
         print*, "Local time stepping is not offloaded to the GPU." ; errorMessage(STD_OUT) ; error stop
      end if
      call check(h3d_max_timestep(h3d, cfl, dcfl, dt_conv, dt_visc), "h3d_max_timestep")
      MaxDt = min(dt_conv, dt_visc)                                          ! :1025-1031
   end subroutine MaxTimeStep_GPU

   function ScalarVolumeIntegral_GPU(integralType) result(val)              ! VolumeIntegrals.f90:76-120 (H3D_INT_* kinds)
      integer, intent(in) :: integralType
      real(kind=RP)       :: val
      call check(h3d_volume_integral(h3d, int(integralType, c_int), val), "h3d_volume_integral")
   end function ScalarVolumeIntegral_GPU

   function VectorSurfaceIntegral_GPU(zoneID, integralType) result(F)       ! SurfaceIntegrals.f90:40, 248 (H3D_SURF_* kinds)
      integer, intent(in) :: zoneID, integralType
      real(kind=RP)       :: F(NDIM)
      call check(h3d_surface_integral(h3d, int(zoneID - 1, c_int), int(integralType, c_int), F), "h3d_surface_integral")
   end function VectorSurfaceIntegral_GPU

   subroutine checkForNan_GPU(mesh, t)                                      ! ExplicitMethods.f90:1856-1905
      type(HexMesh), target, intent(inout) :: mesh
      real(kind=RP),         intent(in)    :: t
      integer(c_int) :: flag
      call check(h3d_has_nan(h3d, flag), "h3d_has_nan")
      if ( flag /= 0 ) then
         call h3d_gpu_download_state(mesh, .false.)                          ! the reference saves the diverged field, then exits
         print*, "Numerical divergence obtained in solver."
         call exit(99)
      end if
   end subroutine checkForNan_GPU

end module H3DGpuAdapter
