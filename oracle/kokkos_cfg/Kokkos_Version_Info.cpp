// Stand-in for the CMake-generated Kokkos_Version_Info.cpp (oracle build only).
#include "Kokkos_Version_Info.hpp"
namespace Kokkos {
namespace Impl {
std::string GIT_BRANCH             = "vendored";
std::string GIT_COMMIT_HASH        = "1a3ea28";
std::string GIT_CLEAN_STATUS       = "CLEAN";
std::string GIT_COMMIT_DESCRIPTION = "Kokkos 4.1.00 as vendored by the reference";
std::string GIT_COMMIT_DATE        = "";
}  // namespace Impl
}  // namespace Kokkos
