!> ISO_C_BINDING shims for builds that prefer explicit interfaces over relying on gfortran's
!> symbol mangling.  UNTESTED IN THIS REPOSITORY'S ENVIRONMENT: no Fortran compiler exists in the
!> build image (see DESIGN.md section 1).  They forward to the neutral C entry points of
!> libdsurf_b200.so (include/dsurftomo_b200.h) and reproduce the reference's STOP behaviour.
!>
!> Simplest integration needs NO shim at all: libdsurf_b200.so already exports the gfortran
!> symbols calsurfg_, depthkernel_, caldespersion_, surfdisp96_, aprod_ and
!> __lsmrmodule_MOD_lsmr, so src/main.f90 links against it unchanged once CalSurfG.o,
!> surfdisp96.o, aprod.o and lsmrModule.o are dropped from the link line (the lsmrmodule.mod
!> file produced by compiling the original lsmrModule.f90 is still used at compile time).
module dsurf_b200_c
  use iso_c_binding
  implicit none
  interface
    integer(c_int) function dsurf_lsmr(m, n, leniw, lenrw, iw, rw, b, damp, atol, btol, conlim, &
        itnlim, localSize, x, istop, itn, normA, condA, normr, normAr, normx) bind(C, name="dsurf_lsmr")
      import :: c_int, c_float
      integer(c_int), value :: m, n, leniw, lenrw, itnlim, localSize
      integer(c_int), intent(in) :: iw(*)
      real(c_float), intent(in) :: rw(*), b(*)
      real(c_float), value :: damp, atol, btol, conlim
      real(c_float), intent(out) :: x(*)
      integer(c_int), intent(out) :: istop, itn
      real(c_float), intent(out) :: normA, condA, normr, normAr, normx
    end function
    integer(c_int) function dsurf_surfdisp96(thkm, vpm, vsm, rhom, nlayer, iflsph, iwave, mode, igr, &
        kmax, t, cg) bind(C, name="dsurf_surfdisp96")
      import :: c_int, c_float, c_double
      real(c_float), intent(in) :: thkm(*), vpm(*), vsm(*), rhom(*)
      integer(c_int), value :: nlayer, iflsph, iwave, mode, igr, kmax
      real(c_double), intent(in) :: t(*)
      real(c_double), intent(out) :: cg(*)
    end function
    subroutine dsurf_fatal(rc) bind(C, name="dsurf_fatal_")
      import :: c_int
      integer(c_int), intent(in) :: rc
    end subroutine
  end interface
end module dsurf_b200_c

!> Drop-in replacement of module LSMRmodule (src/lsmrModule.f90:36-38): same module name, same
!> public routine, same argument list; forwards to the GPU solver.
module LSMRmodule
  use iso_c_binding
  use dsurf_b200_c
  implicit none
  private
  public :: LSMR
contains
  subroutine LSMR(m, n, leniw, lenrw, iw, rw, b, damp, atol, btol, conlim, itnlim, localSize, nout, &
                  x, istop, itn, normA, condA, normr, normAr, normx)
    integer, intent(in) :: leniw, lenrw
    integer, intent(in) :: iw(leniw)
    real, intent(in) :: rw(lenrw)
    integer, intent(in) :: m, n, itnlim, localSize, nout
    integer, intent(out) :: istop, itn
    real(4), intent(in) :: b(m)
    real(4), intent(out) :: x(n)
    real(4), intent(in) :: atol, btol, conlim, damp
    real(4), intent(out) :: normA, condA, normr, normAr, normx
    integer(c_int) :: rc
    rc = dsurf_lsmr(m, n, leniw, lenrw, iw, rw, b, damp, atol, btol, conlim, itnlim, localSize, &
                    x, istop, itn, normA, condA, normr, normAr, normx)
    if (rc /= 0) call dsurf_fatal(rc)
  end subroutine LSMR
end module LSMRmodule
