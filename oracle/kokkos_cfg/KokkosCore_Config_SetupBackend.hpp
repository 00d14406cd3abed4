#ifndef KOKKOS_SETUP_HPP_
#define KOKKOS_SETUP_HPP_
/* host-only build: no device backend setup header */
#endif
