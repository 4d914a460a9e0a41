#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Writes the feature-switch headers that the reference's cmake
# configure step would generate (Kokkos 5.1.0: OpenMP+Serial host backends, no hwloc;
# euler_kokkos: double precision, no MPI, no HDF5).  Usage: gen_config.sh <gen-dir>
set -e
G="$1"
mkdir -p "$G/src/shared" "$G/desul/atomics"

cat > "$G/KokkosCore_config.h" <<'EOF'
#if !defined(KOKKOS_MACROS_HPP) || defined(KOKKOS_CORE_CONFIG_H)
#error "Do not include KokkosCore_config.h directly; include Kokkos_Macros.hpp instead."
#else
#define KOKKOS_CORE_CONFIG_H
#endif
#define KOKKOS_VERSION 50100
#define KOKKOS_VERSION_MAJOR 5
#define KOKKOS_VERSION_MINOR 1
#define KOKKOS_VERSION_PATCH 0
#define KOKKOS_ENABLE_SERIAL
#define KOKKOS_ENABLE_OPENMP
#define KOKKOS_ENABLE_CXX20
#define KOKKOS_ENABLE_DEPRECATED_CODE_5
#define KOKKOS_ENABLE_IMPL_MDSPAN
#define KOKKOS_ENABLE_IMPL_REF_COUNT_BRANCH_UNLIKELY
#define KOKKOS_ENABLE_LIBDL
#define KOKKOS_IMPL_DESUL_VERSION "unknown"
#define KOKKOS_IMPL_MDSPAN_VERSION "unknown"
EOF

cat > "$G/KokkosCore_Config_FwdBackend.hpp" <<'EOF'
#ifndef KOKKOS_FWD_HPP_
#define KOKKOS_FWD_HPP_
#include <fwd/Kokkos_Fwd_SERIAL.hpp>
#include <fwd/Kokkos_Fwd_OPENMP.hpp>
#endif
EOF
cat > "$G/KokkosCore_Config_SetupBackend.hpp" <<'EOF'
#ifndef KOKKOS_SETUP_HPP_
#define KOKKOS_SETUP_HPP_
#endif
EOF
cat > "$G/KokkosCore_Config_DeclareBackend.hpp" <<'EOF'
#ifndef KOKKOS_DECLARE_HPP_
#define KOKKOS_DECLARE_HPP_
#include <decl/Kokkos_Declare_SERIAL.hpp>
#include <decl/Kokkos_Declare_OPENMP.hpp>
#endif
EOF
cat > "$G/KokkosCore_Config_PostInclude.hpp" <<'EOF'
#ifndef KOKKOS_POST_INCLUDE_HPP_
#define KOKKOS_POST_INCLUDE_HPP_
#endif
EOF

cat > "$G/desul/atomics/Config.hpp" <<'EOF'
#ifndef DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_ENABLE_OPENMP
#endif
EOF

cat > "$G/Kokkos_Version_Info.cpp" <<'EOF'
#include "Kokkos_Version_Info.hpp"
namespace Kokkos { namespace Impl {
std::string GIT_BRANCH = "";
std::string GIT_COMMIT_HASH = "";
std::string GIT_CLEAN_STATUS = "";
std::string GIT_COMMIT_DESCRIPTION = "";
std::string GIT_COMMIT_DATE = "";
} }
EOF

cat > "$G/src/shared/euler_kokkos_config.h" <<'EOF'
#ifndef EULER_KOKKOS_CONFIG_H
#define EULER_KOKKOS_CONFIG_H 1
#define EULER_KOKKOS_USE_DOUBLE 1
#define EULER_KOKKOS_TIMING_ENABLED 1
#define EULER_KOKKOS_BUILD_TYPE "Release"
#define EULER_KOKKOS_BUILD_COMPILER_ID "GNU"
#define EULER_KOKKOS_BUILD_COMPILER_VERSION ""
#define EULER_KOKKOS_SYSTEM_PROCESSOR "x86_64"
#define EULER_KOKKOS_SYSTEM_NAME "Linux"
#define EULER_KOKKOS_MPI_VENDOR ""
#endif
EOF

cat > "$G/src/euler_kokkos_version.h" <<'EOF'
#ifndef EULER_KOKKOS_VERSION_H_
#define EULER_KOKKOS_VERSION_H_
#define EULER_KOKKOS_VERSION "ref"
#define EULER_KOKKOS_GIT_TAG ""
#define EULER_KOKKOS_GIT_HASH "62985b2c"
#define EULER_KOKKOS_GIT_HEAD ""
#define EULER_KOKKOS_GIT_HAS_LOCAL_CHANGES ""
#define EULER_KOKKOS_GIT_IS_CLEAN 1
#define EULER_KOKKOS_GIT_REMOTE_URL ""
#define EULER_KOKKOS_GIT_BRANCH ""
#define EULER_KOKKOS_COMPILE_DATE ""
#define EULER_KOKKOS_COMPILE_TIME ""
#endif
EOF
