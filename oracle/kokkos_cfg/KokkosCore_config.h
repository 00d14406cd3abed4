/* Hand-written stand-in for the header Kokkos' CMake would generate
 * (template: external/kokkos/cmake/KokkosCore_config.h.in).  Test infrastructure
 * only: it lets oracle/Makefile compile the vendored Kokkos 4.1.00 OpenMP+Serial
 * host backends with plain g++ so the UNMODIFIED reference headers can be run as
 * the parity oracle.  Nothing here is part of the product. */
#if !defined(KOKKOS_MACROS_HPP) || defined(KOKKOS_CORE_CONFIG_H)
#error "Do not include KokkosCore_config.h directly; include Kokkos_Macros.hpp instead."
#else
#define KOKKOS_CORE_CONFIG_H
#endif
#define KOKKOS_VERSION 40100
#define KOKKOS_VERSION_MAJOR 4
#define KOKKOS_VERSION_MINOR 1
#define KOKKOS_VERSION_PATCH 0
#define KOKKOS_ENABLE_SERIAL
#define KOKKOS_ENABLE_OPENMP
#define KOKKOS_ENABLE_CXX20
#define KOKKOS_ENABLE_DEPRECATED_CODE_4
#define KOKKOS_ENABLE_LIBDL
