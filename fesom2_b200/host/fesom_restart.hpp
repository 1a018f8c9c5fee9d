// fesom_restart.hpp -- read_all_bin_restarts for the C++ host side: the derived-type binary restarts an unmodified FESOM2
// writes with write_all_bin_restarts and the reference's dwarfs read (src/io_restart_derivedtype.F90:29-234): t_mesh.<rank>,
// t_partit.<rank>, t_tracer.<rank>, t_dynamics.<rank>, one Fortran sequential unformatted record each (gfortran markers, long
// records split into sub-records), the items in the order of the WRITE_T_* procedures (src/MOD_MESH.F90:179-279,
// src/MOD_PARTIT.F90:123-193, src/MOD_TRACER.F90:112-177, src/MOD_DYN.F90:212-340).  Same field lists as
// fesom2_b200/restart.py, whose schemas tests/test_restart_io.py checks against the reference sources.
#pragma once
#include <string>

#include "fesom_host.hpp"

namespace fesom {

// mpirank_to_txt (src/fortran_utils.F90:45-56): the rank padded to the width of npes
std::string mpirank_to_txt(int mype, int npes);

// read_all_bin_restarts(path_in, partit, mesh, dynamics, tracers) (src/io_restart_derivedtype.F90:153-234).  `fer_gm` is the
// run's Fer_GM switch, which decides whether fer_w / fer_uv follow cfl_z in t_dynamics (src/MOD_DYN.F90:327-330).
// Throws std::runtime_error with the file name and the byte position when a file does not match the format.
void read_all_bin_restarts(const std::string& path_in, int mype, int npes, t_partit& partit, t_mesh& mesh, t_dyn& dynamics,
                           t_tracer& tracers, bool fer_gm = false);

}  // namespace fesom
