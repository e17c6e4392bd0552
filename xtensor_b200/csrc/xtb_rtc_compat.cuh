// xtb_rtc_compat.cuh -- lets the kernel headers compile both under nvcc (ahead of time) and
// under NVRTC (run-time specialisation of the same templates for arbitrary expression programs,
// xtb_jit.cu).  NVRTC has no host standard library: fixed-width integers and type traits come
// from libcu++ (<cuda/std/...>, shipped with the CUDA toolkit).
#pragma once
#ifdef __CUDACC_RTC__
#include <cuda/std/cstdint>
#include <cuda/std/type_traits>
using cuda::std::int8_t;
using cuda::std::int16_t;
using cuda::std::int32_t;
using cuda::std::int64_t;
using cuda::std::uint8_t;
using cuda::std::uint16_t;
using cuda::std::uint32_t;
using cuda::std::uint64_t;
using cuda::std::uintptr_t;
using cuda::std::size_t;
namespace std {
using cuda::std::bool_constant;
using cuda::std::conditional_t;
using cuda::std::enable_if_t;
using cuda::std::is_floating_point_v;
using cuda::std::is_integral_v;
using cuda::std::is_same_v;
using cuda::std::is_signed_v;
}  // namespace std
#ifndef NAN
#define NAN __int_as_float(0x7fc00000)
#endif
#define XTB_RTC 1
#else
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
#endif
