/* oracle/mkl_shim/mkl_types.h -- TEST INFRASTRUCTURE ONLY.
 * Stand-in for Intel MKL's mkl_types.h (MKL is a third-party dependency that is
 * not installed in this image and is not part of /root/reference).  It lets the
 * UNMODIFIED reference sources compile; see oracle/Makefile. LP64: MKL_INT = int
 * (reference links mkl_intel_lp64, CMakeLists.txt:24). */
#pragma once
#ifndef MKL_INT
#define MKL_INT int
#endif
