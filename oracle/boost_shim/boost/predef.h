/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <boost/predef.h>, see predef/version_number.h.
 * Supports exactly one toolchain: g++ on Linux / x86-64 (what this image has).  Every detection macro alpaka 1.2.0
 * tests is defined; the ones that do not apply are BOOST_VERSION_NUMBER_NOT_AVAILABLE (= 0), as in Boost.Predef. */
#ifndef PPS_BOOST_SHIM_PREDEF_H
#define PPS_BOOST_SHIM_PREDEF_H

#include <boost/predef/version_number.h>

#if !defined(__GNUC__) || defined(__clang__) || defined(__CUDACC__) || !defined(__linux__)
#    error "oracle/boost_shim supports g++ on Linux only"
#endif

/* compilers */
#define BOOST_COMP_GNUC BOOST_VERSION_NUMBER(__GNUC__, __GNUC_MINOR__, __GNUC_PATCHLEVEL__)
#define BOOST_COMP_GNUC_AVAILABLE
#define BOOST_COMP_CLANG BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_NVCC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_MSVC BOOST_VERSION_NUMBER_NOT_AVAILABLE
/* BOOST_COMP_*_EMULATED stay undefined (Boost.Predef defines them only when a front end emulates that compiler) */
#define BOOST_COMP_HPACC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_SUNPRO BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_IBM BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_INTEL BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_PGI BOOST_VERSION_NUMBER_NOT_AVAILABLE

/* languages */
#define BOOST_LANG_CUDA BOOST_VERSION_NUMBER_NOT_AVAILABLE

/* operating systems */
#define BOOST_OS_LINUX BOOST_VERSION_NUMBER_AVAILABLE
#define BOOST_OS_UNIX BOOST_VERSION_NUMBER_AVAILABLE
#define BOOST_OS_WINDOWS BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_MACOS BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_CYGWIN BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_BSD BOOST_VERSION_NUMBER_NOT_AVAILABLE

/* architectures */
#if defined(__x86_64__) || defined(__i386__)
#    define BOOST_ARCH_X86 BOOST_VERSION_NUMBER_AVAILABLE
#else
#    define BOOST_ARCH_X86 BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#define BOOST_ARCH_PTX BOOST_VERSION_NUMBER_NOT_AVAILABLE

#endif
