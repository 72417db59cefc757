/* Compile-time constants (the reference generates these from CMake, reference CMakeLists.txt:13-21). */
#ifndef LIBNOMP_B200_DEFS_H_
#define LIBNOMP_B200_DEFS_H_

#define NOMP_MAX_BUFFER_SIZE 128
#define NOMP_MAX_KERNEL_ARGS_SIZE 64
#define NOMP_MAX_SCRATCH_SIZE 131072 /* doubles: reduction workspace of libnompk (536 KiB) + result slot */

#define NOMP_DEFAULT_VERBOSE 2
#define NOMP_DEFAULT_PROFILE 0
#define NOMP_DEFAULT_DEVICE 0
#define NOMP_DEFAULT_PLATFORM 0

#endif
