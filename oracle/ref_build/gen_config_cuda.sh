#!/bin/sh
# TEST / BENCH INFRASTRUCTURE ONLY.  Feature-switch headers for the reference's OTHER build: Kokkos 5.1.0 with the
# Cuda backend (+ Serial host backend) for NVIDIA Blackwell CC 10.0 -- what `cmake -DEULER_KOKKOS_BACKEND=Cuda
# -DKokkos_ARCH_BLACKWELL100=ON` would configure.  Usage: gen_config_cuda.sh <gen-dir>
set -e
G="$1"
sh "$(dirname "$0")/gen_config.sh" "$G"

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
#define KOKKOS_ENABLE_CUDA
#define KOKKOS_COMPILER_CUDA_VERSION 129
#define KOKKOS_ENABLE_CXX20
#define KOKKOS_ENABLE_CUDA_LAMBDA
#define KOKKOS_ENABLE_CUDA_CONSTEXPR
#define KOKKOS_ENABLE_IMPL_CUDA_MALLOC_ASYNC
#define KOKKOS_ENABLE_DEPRECATED_CODE_5
#define KOKKOS_ENABLE_IMPL_MDSPAN
#define KOKKOS_ENABLE_IMPL_REF_COUNT_BRANCH_UNLIKELY
#define KOKKOS_ENABLE_LIBDL
#define KOKKOS_ARCH_BLACKWELL
#define KOKKOS_ARCH_BLACKWELL100
#define KOKKOS_IMPL_DESUL_VERSION "unknown"
#define KOKKOS_IMPL_MDSPAN_VERSION "unknown"
EOF

cat > "$G/KokkosCore_Config_FwdBackend.hpp" <<'EOF'
#ifndef KOKKOS_FWD_HPP_
#define KOKKOS_FWD_HPP_
#include <fwd/Kokkos_Fwd_SERIAL.hpp>
#include <fwd/Kokkos_Fwd_CUDA.hpp>
#endif
EOF
cat > "$G/KokkosCore_Config_SetupBackend.hpp" <<'EOF'
#ifndef KOKKOS_SETUP_HPP_
#define KOKKOS_SETUP_HPP_
#include <setup/Kokkos_Setup_Cuda.hpp>
#endif
EOF
cat > "$G/KokkosCore_Config_DeclareBackend.hpp" <<'EOF'
#ifndef KOKKOS_DECLARE_HPP_
#define KOKKOS_DECLARE_HPP_
#include <decl/Kokkos_Declare_SERIAL.hpp>
#include <decl/Kokkos_Declare_CUDA.hpp>
#endif
EOF

cat > "$G/desul/atomics/Config.hpp" <<'EOF'
#ifndef DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_CONFIG_HPP_
#define DESUL_ATOMICS_ENABLE_CUDA
#endif
EOF
