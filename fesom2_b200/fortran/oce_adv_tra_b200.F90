!===============================================================================
! oce_adv_tra_b200.F90 -- the thin ISO_C_BINDING layer between an unmodified FESOM2 host and
! libfesom_adv_b200.so (include/fesom_adv_b200.h).
!
! Drop-in for the reference's external procedure
!     subroutine do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh)
! (src/oce_adv_tra_driver.F90:46-490, interface module :1-21): same name, same argument list, same
! in-place accumulation into tracers%work%del_ttf_advhoriz / del_ttf_advvert.  Link this file INSTEAD
! of src/oce_adv_tra_driver.F90 (keep src/oce_adv_tra_{hor,ver,fct}.F90 out of the link as well: the
! library replaces them) and add -lfesom_adv_b200 to the link line.  Nothing else in FESOM changes:
! solve_tracers_ale (src/oce_ale_tracer.F90:312) and the dwarf (dwarf_ini/fesom.F90:97) call it as
! before.
!
! Derived types with allocatable components are not C-interoperable, so this wrapper passes the
! ADDRESS of every component the path reads plus the scalar dimensions.  Which address is decided in
! ONE place, the macro ADV_ADDR: c_loc(x) in a host build, acc_deviceptr(x) (OpenACC >= 2.6 Fortran
! API, module openacc) in the reference's GPU build, where the operands live in the enclosing
! `!$ACC DATA` / `!$ACC ENTER DATA` regions (dwarf_ini/fesom.F90:71-83, src/fesom_module.F90:759-805).
! acc_deviceptr is an ordinary function: unlike `!$ACC HOST_DATA USE_DEVICE`, which is lexical to the
! region it is written in, it returns the device address in any scope.
!
! State upload: the reference refreshes uv/w/thicknesses once per step (src/oce_ale_tracer.F90:260-262)
! and then calls do_oce_adv_tra once per tracer (:280-312).  The wrapper hands the state over with
! adv_ctx_set_state_step and the model's step counter mstep (src/oce_modules.F90:23): only the first
! tracer of a step uploads it and computes the edge volume flux, the others reuse both.
!
! Batched entry: do_oce_adv_tra_b200_batch hands tracers tr_first..tr_last to ONE library call so that
! geometry reads are amortised (RECOM-style runs).  The reference owns ONE tracers%work set
! (edge_up_dn_grad, del_ttf_advhoriz, del_ttf_advvert: src/MOD_TRACER.F90:36,64), which cannot serve
! several tracers at once: a batch of more than one tracer must bring per-tracer arrays (the optional
! arguments grad_b, dh_b, dv_b); without them the call is refused.
!
! This image has no Fortran compiler: the file is shipped as source and compiled by the FESOM build
! (INTEGRATION.md); tests/test_fortran_shim.py checks every bind(C) type and interface of this file
! field by field against include/fesom_adv_b200.h.
!===============================================================================
#ifdef ENABLE_OPENACC
#define ADV_ADDR(x) acc_deviceptr(x)
#define ADV_WHERE ADV_DEVICE
#else
#define ADV_ADDR(x) c_loc(x)
#define ADV_WHERE ADV_HOST
#endif

module oce_adv_tra_b200
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: adv_b200_init, adv_b200_finalize, do_oce_adv_tra_b200_batch, adv_b200_ctx

  integer(c_int), parameter :: ADV_OK = 0, ADV_EINVAL = -1, ADV_ESCHEME = -3
  integer(c_int), parameter :: ADV_HOST = 0, ADV_DEVICE = 1

  ! adv_mesh_desc_t (include/fesom_adv_b200.h)
  type, bind(C) :: adv_mesh_desc_t
     integer(c_int32_t) :: nl
     integer(c_int32_t) :: myDim_nod2D, eDim_nod2D
     integer(c_int32_t) :: myDim_elem2D, eDim_elem2D
     integer(c_int32_t) :: myDim_edge2D
     integer(c_int32_t) :: nod_in_elem2D_ld
     type(c_ptr) :: edges, edge_tri, elem2D_nodes, nod_in_elem2D, nod_in_elem2D_num
     type(c_ptr) :: nlevels, ulevels, nlevels_nod2D, ulevels_nod2D
     type(c_ptr) :: edge_cross_dxdy, edge_dxdy, elem_cos, area, areasvol, nboundary_lay
     integer(c_int32_t) :: mype, npes
     integer(c_int32_t) :: rPEnum
     type(c_ptr) :: rPE, rptr, rlist
     integer(c_int32_t) :: sPEnum
     type(c_ptr) :: sPE, sptr, slist
  end type adv_mesh_desc_t

  ! adv_state_desc_t
  type, bind(C) :: adv_state_desc_t
     type(c_ptr) :: uv, w, w_e, w_i, helem, hnode, hnode_new, zbar_3d_n, Z_3d_n, zbar_n_bot
     integer(c_int32_t) :: use_wsplit
  end type adv_state_desc_t

  ! adv_tracer_desc_t
  type, bind(C) :: adv_tracer_desc_t
     type(c_ptr) :: values, valuesAB, edge_up_dn_grad, del_ttf_advhoriz, del_ttf_advvert
     type(c_ptr) :: tra_adv_hor, tra_adv_ver, tra_adv_lim
     real(c_double) :: tra_adv_ph, tra_adv_pv
     type(c_ptr) :: tra_advhoriz, tra_advvert
     type(c_ptr) :: dvd_trflx_hor, dvd_trflx_ver
  end type adv_tracer_desc_t

  ! adv_gradient_mesh_desc_t
  type, bind(C) :: adv_gradient_mesh_desc_t
     integer(c_int32_t) :: n_elem
     integer(c_int32_t) :: n_nod_in_elem
     integer(c_int32_t) :: nod_in_elem2D_ld
     type(c_ptr) :: nod_in_elem2D, nod_in_elem2D_num
     type(c_ptr) :: nlevels, ulevels
     type(c_ptr) :: edge_up_dn_tri, nlevels_nod2D_min, ulevels_nod2D_max
     type(c_ptr) :: gradient_sca, elem_area
     integer(c_int32_t) :: rPEnum
     type(c_ptr) :: rPE, rptr, rlist
     integer(c_int32_t) :: sPEnum
     type(c_ptr) :: sPE, sptr, slist
  end type adv_gradient_mesh_desc_t

  ! adv_zstar_desc_t
  type, bind(C) :: adv_zstar_desc_t
     type(c_ptr) :: hbar, hbar_old
     type(c_ptr) :: water_flux
     type(c_ptr) :: nlevels_nod2D_min
     type(c_ptr) :: hnode_new
  end type adv_zstar_desc_t

  ! adv_zlevel_desc_t
  type, bind(C) :: adv_zlevel_desc_t
     type(c_ptr) :: hbar, hbar_old
     type(c_ptr) :: water_flux
     type(c_ptr) :: nlevels_nod2D_min
     type(c_ptr) :: hnode_new
     type(c_ptr) :: zbar
     real(c_double) :: min_hnode
     integer(c_int32_t) :: lzstar_lev
  end type adv_zlevel_desc_t

  interface
     integer(c_int) function adv_ctx_create(ctx, mesh, device, max_tracers) bind(C, name='adv_ctx_create')
       import :: c_int, c_ptr, adv_mesh_desc_t
       type(c_ptr), intent(out) :: ctx
       type(adv_mesh_desc_t), intent(in) :: mesh
       integer(c_int), value :: device, max_tracers
     end function
     integer(c_int) function adv_ctx_destroy(ctx) bind(C, name='adv_ctx_destroy')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     type(c_ptr) function adv_last_error() bind(C, name='adv_last_error')
       import :: c_ptr
     end function
     integer(c_int) function adv_comm_unique_id(id) bind(C, name='adv_comm_unique_id')
       import :: c_int, c_char
       character(kind=c_char) :: id(128)
     end function
     integer(c_int) function adv_ctx_comm_init(ctx, id) bind(C, name='adv_ctx_comm_init')
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char) :: id(128)
     end function
     integer(c_int) function adv_ctx_set_state(ctx, st, where) bind(C, name='adv_ctx_set_state')
       import :: c_int, c_ptr, adv_state_desc_t
       type(c_ptr), value :: ctx
       type(adv_state_desc_t), intent(in) :: st
       integer(c_int), value :: where
     end function
     integer(c_int) function adv_ctx_set_state_step(ctx, st, where, step) bind(C, name='adv_ctx_set_state_step')
       import :: c_int, c_int64_t, c_ptr, adv_state_desc_t
       type(c_ptr), value :: ctx
       type(adv_state_desc_t), intent(in) :: st
       integer(c_int), value :: where
       integer(c_int64_t), value :: step
     end function
     integer(c_int) function adv_do_oce_adv_tra(ctx, dt, ntr, tr, where) bind(C, name='adv_do_oce_adv_tra')
       import :: c_int, c_ptr, c_double, adv_tracer_desc_t
       type(c_ptr), value :: ctx
       real(c_double), value :: dt
       integer(c_int), value :: ntr
       type(adv_tracer_desc_t), intent(in) :: tr(*)
       integer(c_int), value :: where
     end function
     ! the producer of edge_up_dn_grad on the device (tracer_gradient_elements, exchange_elem, fill_up_dn_grad)
     integer(c_int) function adv_ctx_set_gradient_mesh(ctx, g) bind(C, name='adv_ctx_set_gradient_mesh')
       import :: c_ptr, c_int, adv_gradient_mesh_desc_t
       type(c_ptr), value :: ctx
       type(adv_gradient_mesh_desc_t), intent(in) :: g
     end function
     integer(c_int) function adv_tracer_gradient_elements(ctx, ntr, ttf, tr_xy) bind(C, name='adv_tracer_gradient_elements')
       import :: c_ptr, c_int
       type(c_ptr), value    :: ctx
       integer(c_int), value :: ntr
       type(c_ptr), intent(in) :: ttf(*), tr_xy(*)            ! arrays of ntr device pointers
     end function
     integer(c_int) function adv_fill_up_dn_grad(ctx, ntr, tr_xy, edge_up_dn_grad) bind(C, name='adv_fill_up_dn_grad')
       import :: c_ptr, c_int
       type(c_ptr), value    :: ctx
       integer(c_int), value :: ntr
       type(c_ptr), intent(in) :: tr_xy(*), edge_up_dn_grad(*)
     end function
     integer(c_int) function adv_exchange_elem(ctx, nfields, fields, nwords) bind(C, name='adv_exchange_elem')
       import :: c_ptr, c_int
       type(c_ptr), value    :: ctx
       integer(c_int), value :: nfields
       type(c_ptr), intent(in) :: fields(*)
       integer(c_int), value :: nwords
     end function
     ! device-resident dwarf loop (dwarf_ini/fesom.F90:85-128): prologue, epilogue, halo exchange
     integer(c_int) function adv_init_tracers_AB(ctx, ntr, ab_order, epsilon, values, valuesold, valuesAB, &
                                                 del_ttf, del_ttf_advhoriz, del_ttf_advvert) bind(C, name='adv_init_tracers_AB')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: ntr, ab_order
       real(c_double), value :: epsilon
       type(c_ptr), intent(in) :: values(*), valuesold(*), valuesAB(*), del_ttf(*), del_ttf_advhoriz(*), del_ttf_advvert(*)
     end function
     integer(c_int) function adv_update_values(ctx, ntr, values, del_ttf_advhoriz, del_ttf_advvert) bind(C, name='adv_update_values')
       import :: c_ptr, c_int
       type(c_ptr), value    :: ctx
       integer(c_int), value :: ntr
       type(c_ptr), intent(in) :: values(*), del_ttf_advhoriz(*), del_ttf_advvert(*)
     end function
     integer(c_int) function adv_exchange_nod(ctx, nfields, fields, nlev) bind(C, name='adv_exchange_nod')
       import :: c_ptr, c_int
       type(c_ptr), value    :: ctx
       integer(c_int), value :: nfields
       type(c_ptr), intent(in) :: fields(*)
       integer(c_int), value :: nlev
     end function
     ! vert_vel_ale (continuity part, exchange, compute_CFLz, compute_Wvel_split) for linfs / zstar, src/oce_ale.F90:2107-2668
     integer(c_int) function adv_vert_vel_ale(ctx, dt, use_wsplit, wsplit_maxcfl, w, w_e, w_i, cfl_z) bind(C, name='adv_vert_vel_ale')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: dt
       integer(c_int), value :: use_wsplit
       real(c_double), value :: wsplit_maxcfl
       type(c_ptr), value    :: w, w_e, w_i, cfl_z
     end function
     integer(c_int) function adv_vert_vel_ale_zstar(ctx, dt, use_wsplit, wsplit_maxcfl, z, w, w_e, w_i, cfl_z) &
                                                    bind(C, name='adv_vert_vel_ale_zstar')
       import :: c_ptr, c_int, c_double, adv_zstar_desc_t
       type(c_ptr), value    :: ctx
       real(c_double), value :: dt
       integer(c_int), value :: use_wsplit
       real(c_double), value :: wsplit_maxcfl
       type(adv_zstar_desc_t), intent(in) :: z
       type(c_ptr), value    :: w, w_e, w_i, cfl_z
     end function
     ! which_ALE = 'zlevel' (src/oce_ale.F90:2336-2538); cfl_z is in/out (the previous step's CFL_z on entry)
     integer(c_int) function adv_vert_vel_ale_zlevel(ctx, dt, use_wsplit, wsplit_maxcfl, z, w, w_e, w_i, cfl_z) &
                                                     bind(C, name='adv_vert_vel_ale_zlevel')
       import :: c_ptr, c_int, c_double, adv_zlevel_desc_t
       type(c_ptr), value    :: ctx
       real(c_double), value :: dt
       integer(c_int), value :: use_wsplit
       real(c_double), value :: wsplit_maxcfl
       type(adv_zlevel_desc_t), intent(in) :: z
       type(c_ptr), value    :: w, w_e, w_i, cfl_z
     end function
     integer(c_int) function adv_ctx_wait_for(ctx, stream) bind(C, name='adv_ctx_wait_for')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx, stream
     end function
     integer(c_int) function adv_ctx_signal(ctx, stream) bind(C, name='adv_ctx_signal')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx, stream
     end function
     integer(c_int) function adv_ctx_synchronize(ctx) bind(C, name='adv_ctx_synchronize')
       import :: c_ptr, c_int
       type(c_ptr), value :: ctx
     end function
  end interface

  type(c_ptr), save :: adv_b200_ctx = c_null_ptr   ! one context per MPI rank (one rank <-> one GPU)
  ! .true.: the library computes edge_up_dn_grad itself (tracer_gradient_elements, exchange_elem, fill_up_dn_grad on
  ! the device; the wrapper then passes a NULL gradient pointer and the 4 E (nl-1) words per tracer never cross PCIe)
  logical, save :: adv_b200_device_gradients = .false.

contains

  !-----------------------------------------------------------------------------
  ! Call once after oce_adv_tra_fct_init / muscl_adv_init (src/oce_setup_step.F90:239): uploads the
  ! static mesh slice, builds the gather lists and the NCCL communicator.  The descriptor's arrays are
  ! always HOST arrays (the library copies them).
  !-----------------------------------------------------------------------------
  subroutine adv_b200_init(tracers, partit, mesh, device, device_gradients)
    use MOD_MESH
    use MOD_TRACER
    use MOD_PARTIT
    use MOD_PARSUP
    use mpi
    type(t_tracer), intent(inout), target :: tracers
    type(t_partit), intent(inout), target :: partit
    type(t_mesh),   intent(in),    target :: mesh
    integer,        intent(in)            :: device
    logical,        intent(in), optional  :: device_gradients
    type(adv_mesh_desc_t) :: d
    type(adv_gradient_mesh_desc_t) :: gd
    character(kind=c_char) :: id(128)
    integer :: rc, ierr

    d%nl = mesh%nl
    d%myDim_nod2D = partit%myDim_nod2D;   d%eDim_nod2D = partit%eDim_nod2D
    d%myDim_elem2D = partit%myDim_elem2D; d%eDim_elem2D = partit%eDim_elem2D
    d%myDim_edge2D = partit%myDim_edge2D
    d%nod_in_elem2D_ld = size(mesh%nod_in_elem2D, 1)
    d%edges = c_loc(mesh%edges);                 d%edge_tri = c_loc(mesh%edge_tri)
    d%elem2D_nodes = c_loc(mesh%elem2D_nodes);   d%nod_in_elem2D = c_loc(mesh%nod_in_elem2D)
    d%nod_in_elem2D_num = c_loc(mesh%nod_in_elem2D_num)
    d%nlevels = c_loc(mesh%nlevels);             d%ulevels = c_loc(mesh%ulevels)
    d%nlevels_nod2D = c_loc(mesh%nlevels_nod2D); d%ulevels_nod2D = c_loc(mesh%ulevels_nod2D)
    d%edge_cross_dxdy = c_loc(mesh%edge_cross_dxdy); d%edge_dxdy = c_loc(mesh%edge_dxdy)
    d%elem_cos = c_loc(mesh%elem_cos)
    d%area = c_loc(mesh%area);                   d%areasvol = c_loc(mesh%areasvol)
    d%nboundary_lay = c_loc(tracers%work%nboundary_lay)
    d%mype = partit%mype; d%npes = partit%npes
    d%rPEnum = partit%com_nod2D%rPEnum
    d%rPE = c_loc(partit%com_nod2D%rPE); d%rptr = c_loc(partit%com_nod2D%rptr); d%rlist = c_loc(partit%com_nod2D%rlist)
    d%sPEnum = partit%com_nod2D%sPEnum
    d%sPE = c_loc(partit%com_nod2D%sPE); d%sptr = c_loc(partit%com_nod2D%sptr); d%slist = c_loc(partit%com_nod2D%slist)

    rc = adv_ctx_create(adv_b200_ctx, d, int(device, c_int), int(tracers%num_tracers, c_int))
    call check(rc, partit)
    if (partit%npes > 1) then
       if (partit%mype == 0) call check(adv_comm_unique_id(id), partit)
       call MPI_Bcast(id, 128, MPI_BYTE, 0, partit%MPI_COMM_FESOM, ierr)   ! replaces init_mpi_types for this path
       call check(adv_ctx_comm_init(adv_b200_ctx, id), partit)
    end if

    if (present(device_gradients)) adv_b200_device_gradients = device_gradients
    if (adv_b200_device_gradients) then
       ! static inputs of tracer_gradient_elements / fill_up_dn_grad (src/oce_tracer_mod.F90:146-188,
       ! src/oce_muscl_adv.F90:356-525) and the element halo of tr_xy (com_elem2D_full); call after muscl_adv_init
       gd%n_elem = partit%myDim_elem2D + partit%eDim_elem2D + partit%eXDim_elem2D
       gd%n_nod_in_elem = size(mesh%nod_in_elem2D, 2)
       gd%nod_in_elem2D_ld = size(mesh%nod_in_elem2D, 1)
       gd%nod_in_elem2D = c_loc(mesh%nod_in_elem2D); gd%nod_in_elem2D_num = c_loc(mesh%nod_in_elem2D_num)
       gd%nlevels = c_loc(mesh%nlevels);             gd%ulevels = c_loc(mesh%ulevels)
       gd%edge_up_dn_tri = c_loc(tracers%work%edge_up_dn_tri)
       gd%nlevels_nod2D_min = c_loc(mesh%nlevels_nod2D_min); gd%ulevels_nod2D_max = c_loc(mesh%ulevels_nod2D_max)
       gd%gradient_sca = c_loc(mesh%gradient_sca);   gd%elem_area = c_loc(mesh%elem_area)
       gd%rPEnum = partit%com_elem2D_full%rPEnum
       gd%rPE = c_loc(partit%com_elem2D_full%rPE); gd%rptr = c_loc(partit%com_elem2D_full%rptr)
       gd%rlist = c_loc(partit%com_elem2D_full%rlist)
       gd%sPEnum = partit%com_elem2D_full%sPEnum
       gd%sPE = c_loc(partit%com_elem2D_full%sPE); gd%sptr = c_loc(partit%com_elem2D_full%sptr)
       gd%slist = c_loc(partit%com_elem2D_full%slist)
       call check(adv_ctx_set_gradient_mesh(adv_b200_ctx, gd), partit)
    end if
  end subroutine adv_b200_init

  subroutine adv_b200_finalize()
    integer :: rc
    if (c_associated(adv_b200_ctx)) rc = adv_ctx_destroy(adv_b200_ctx)
    adv_b200_ctx = c_null_ptr
  end subroutine adv_b200_finalize

  !-----------------------------------------------------------------------------
  ! tracers tr_first..tr_last in ONE library call.  One tracer: the reference's single work set
  ! tracers%work%{edge_up_dn_grad, del_ttf_advhoriz, del_ttf_advvert} is used, exactly like
  ! do_oce_adv_tra.  More than one: the caller supplies per-tracer arrays
  !     grad_b(4, nl-1, myDim_edge2D, n), dh_b(nl-1, myDim_nod2D+eDim_nod2D, n), dv_b(same)
  ! (n = tr_last-tr_first+1; in an OpenACC build they must be present on the device).
  !-----------------------------------------------------------------------------
  subroutine do_oce_adv_tra_b200_batch(dt, vel, w, wi, we, tr_first, tr_last, dynamics, tracers, partit, mesh, grad_b, dh_b, dv_b)
    use MOD_MESH
    use MOD_TRACER
    use MOD_PARTIT
    use MOD_PARSUP
    use MOD_DYN
    use o_PARAM, only: mstep
    use diagnostics, only: ldiag_DVD
#ifdef ENABLE_OPENACC
    use openacc
#endif
    real(kind=WP),  intent(in),    target :: dt
    integer,        intent(in)            :: tr_first, tr_last
    type(t_partit), intent(inout), target :: partit
    type(t_mesh),   intent(in),    target :: mesh
    type(t_tracer), intent(inout), target :: tracers
    type(t_dyn),    intent(inout), target :: dynamics
    real(kind=WP),  intent(in),    target :: vel(2, mesh%nl-1, partit%myDim_elem2D+partit%eDim_elem2D)
    real(kind=WP),  intent(in),    target :: W(mesh%nl,  partit%myDim_nod2D+partit%eDim_nod2D)
    real(kind=WP),  intent(in),    target :: WI(mesh%nl, partit%myDim_nod2D+partit%eDim_nod2D)
    real(kind=WP),  intent(in),    target :: WE(mesh%nl, partit%myDim_nod2D+partit%eDim_nod2D)
    real(kind=WP),  intent(in),    target, optional :: grad_b(:,:,:,:)
    real(kind=WP),  intent(inout), target, optional :: dh_b(:,:,:), dv_b(:,:,:)
    type(adv_state_desc_t) :: st
    type(adv_tracer_desc_t), allocatable :: td(:)
    character(kind=c_char, len=21), allocatable, target :: hs(:), vs(:), ls(:)
    real(kind=WP), pointer :: helem(:,:), hnode(:,:), hnode_new(:,:), zbar_3d_n(:,:), Z_3d_n(:,:), zbar_n_bot(:)
    real(kind=WP), pointer :: values(:,:), valuesAB(:,:), grad1(:,:,:), dh1(:,:), dv1(:,:), dgh1(:,:), dgv1(:,:)
    integer :: i, k, n, rc

    n = tr_last - tr_first + 1
    if (n > 1 .and. .not. (present(grad_b) .and. present(dh_b) .and. present(dv_b))) then
       if (partit%mype == 0) write(*,*) 'fesom_adv_b200: a batch of ', n, ' tracers needs per-tracer work arrays ', &
                                        '(grad_b, dh_b, dv_b): tracers%work holds one set only (MOD_TRACER.F90:36,64)'
       call par_ex(partit%MPI_COMM_FESOM, partit%mype, 1)
    end if

    helem => mesh%helem; hnode => mesh%hnode; hnode_new => mesh%hnode_new
    zbar_3d_n => mesh%zbar_3d_n; Z_3d_n => mesh%Z_3d_n; zbar_n_bot => mesh%zbar_n_bot
    st%uv = ADV_ADDR(vel); st%w = ADV_ADDR(W); st%w_e = ADV_ADDR(WE); st%w_i = ADV_ADDR(WI)
    st%helem = ADV_ADDR(helem); st%hnode = ADV_ADDR(hnode); st%hnode_new = ADV_ADDR(hnode_new)
    st%zbar_3d_n = ADV_ADDR(zbar_3d_n); st%Z_3d_n = ADV_ADDR(Z_3d_n); st%zbar_n_bot = ADV_ADDR(zbar_n_bot)
    st%use_wsplit = merge(1, 0, dynamics%use_wsplit)
    ! once per model step: a repeated call with the same mstep and the same arrays is a no-op
    call check(adv_ctx_set_state_step(adv_b200_ctx, st, ADV_WHERE, int(mstep, c_int64_t)), partit)

    allocate(td(n), hs(n), vs(n), ls(n))
    do k = 1, n
       i = tr_first + k - 1
       hs(k) = trim(tracers%data(i)%tra_adv_hor)//c_null_char
       vs(k) = trim(tracers%data(i)%tra_adv_ver)//c_null_char
       ls(k) = trim(tracers%data(i)%tra_adv_lim)//c_null_char
       values => tracers%data(i)%values; valuesAB => tracers%data(i)%valuesAB
       td(k)%values   = ADV_ADDR(values)
       td(k)%valuesAB = ADV_ADDR(valuesAB)
       if (n == 1 .and. .not. present(grad_b)) then
          grad1 => tracers%work%edge_up_dn_grad
          dh1 => tracers%work%del_ttf_advhoriz; dv1 => tracers%work%del_ttf_advvert
       else
          grad1 => grad_b(:,:,:,k)                    ! contiguous: the tracer index is the last one
          dh1 => dh_b(:,:,k); dv1 => dv_b(:,:,k)
       end if
       td(k)%edge_up_dn_grad  = ADV_ADDR(grad1)
       if (adv_b200_device_gradients) td(k)%edge_up_dn_grad = c_null_ptr   ! computed by the library from `values`
       td(k)%del_ttf_advhoriz = ADV_ADDR(dh1)
       td(k)%del_ttf_advvert  = ADV_ADDR(dv1)
       td(k)%tra_adv_hor = c_loc(hs(k)); td(k)%tra_adv_ver = c_loc(vs(k)); td(k)%tra_adv_lim = c_loc(ls(k))
       td(k)%tra_adv_ph = tracers%data(i)%tra_adv_ph
       td(k)%tra_adv_pv = tracers%data(i)%tra_adv_pv
       ! ltra_diag (default .true., src/MOD_TRACER.F90:25): the tracer's slices of tracers%work%tra_advhoriz / tra_advvert
       ! (src/oce_adv_tra_driver.F90:221-229, :307-318, :464-488); the tracer index is the last one, so a slice is contiguous
       td(k)%tra_advhoriz = c_null_ptr; td(k)%tra_advvert = c_null_ptr
       if (tracers%data(i)%ltra_diag) then
          dgh1 => tracers%work%tra_advhoriz(:,:,i); dgv1 => tracers%work%tra_advvert(:,:,i)
          td(k)%tra_advhoriz = ADV_ADDR(dgh1); td(k)%tra_advvert = ADV_ADDR(dgv1)
       end if
       ! ldiag_DVD: temperature and salinity only (src/oce_adv_tra_driver.F90:263, :395)
       td(k)%dvd_trflx_hor = c_null_ptr; td(k)%dvd_trflx_ver = c_null_ptr
       if (ldiag_DVD .and. i <= 2) then
          dgh1 => tracers%work%dvd_trflx_hor(:,:,i); dgv1 => tracers%work%dvd_trflx_ver(:,:,i)
          td(k)%dvd_trflx_hor = ADV_ADDR(dgh1); td(k)%dvd_trflx_ver = ADV_ADDR(dgv1)
       end if
    end do
#ifdef ENABLE_OPENACC
    ! the operands may still be in flight on the OpenACC queues: order the library's stream behind them
    !$ACC WAIT
#endif
    rc = adv_do_oce_adv_tra(adv_b200_ctx, real(dt, c_double), int(n, c_int), td, ADV_WHERE)
    call check(rc, partit)
    deallocate(td, hs, vs, ls)
  end subroutine do_oce_adv_tra_b200_batch

  !-----------------------------------------------------------------------------
  ! error convention of the reference: message on rank 0 + par_ex(comm, mype, 1) -> MPI_ABORT
  ! (src/oce_adv_tra_driver.F90:351-353, src/gen_modules_partitioning.F90:87-123)
  !-----------------------------------------------------------------------------
  subroutine check(rc, partit)
    use MOD_PARTIT
    use MOD_PARSUP
    integer(c_int), intent(in) :: rc
    type(t_partit), intent(inout) :: partit
    character(kind=c_char), pointer :: msg(:)
    integer :: k
    if (rc == ADV_OK) return
    call c_f_pointer(adv_last_error(), msg, [512])
    if (partit%mype == 0) then
       do k = 1, 512
          if (msg(k) == c_null_char) exit
       end do
       write(*,*) 'fesom_adv_b200: ', msg(1:k-1)
    end if
    call par_ex(partit%MPI_COMM_FESOM, partit%mype, 1)
  end subroutine check

end module oce_adv_tra_b200

!===============================================================================
! The reference's seam, unchanged: same external procedure name and argument list as
! src/oce_adv_tra_driver.F90:46.  One tracer per call, like the reference; the state upload and the
! volume flux are shared by the calls of one model step (mstep).
!===============================================================================
subroutine do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh)
  use MOD_MESH
  use MOD_TRACER
  use MOD_PARTIT
  use MOD_PARSUP
  use MOD_DYN
  use oce_adv_tra_b200
  implicit none
  real(kind=WP),  intent(in),    target :: dt
  integer,        intent(in)            :: tr_num
  type(t_partit), intent(inout), target :: partit
  type(t_mesh),   intent(in),    target :: mesh
  type(t_tracer), intent(inout), target :: tracers
  type(t_dyn),    intent(inout), target :: dynamics
  real(kind=WP),  intent(in), target    :: vel(2, mesh%nl-1, partit%myDim_elem2D+partit%eDim_elem2D)
  real(kind=WP),  intent(in), target    :: W(mesh%nl,    partit%myDim_nod2D+partit%eDim_nod2D)
  real(kind=WP),  intent(in), target    :: WI(mesh%nl,   partit%myDim_nod2D+partit%eDim_nod2D)
  real(kind=WP),  intent(in), target    :: WE(mesh%nl,   partit%myDim_nod2D+partit%eDim_nod2D)
  call do_oce_adv_tra_b200_batch(dt, vel, W, WI, WE, tr_num, tr_num, dynamics, tracers, partit, mesh)
end subroutine do_oce_adv_tra
