#ifndef KOKKOS_DECLARE_HPP_
#define KOKKOS_DECLARE_HPP_
#include <decl/Kokkos_Declare_SERIAL.hpp>
#include <decl/Kokkos_Declare_OPENMP.hpp>
#endif
