/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for Boost.Predef's version-number macros.
 *
 * This image has no Boost.  The alpaka 1.2.0 headers vendored by the reference
 * (solverPoissonMPI_alpaka/thirdParty/alpaka) need Boost only for compiler / OS / architecture detection
 * (boost/predef.h) and for type-name demangling (boost/core/demangle.hpp); this directory supplies exactly
 * those for a g++ / Linux / x86-64 host build, so that oracle/build_ref_alpaka.py can compile the reference's
 * unmodified alpaka tree with its CPU accelerator.  The encoding follows Boost.Predef's documented scheme
 * (major * 10^7 + minor * 10^5 + patch). */
#ifndef PPS_BOOST_SHIM_VERSION_NUMBER_H
#define PPS_BOOST_SHIM_VERSION_NUMBER_H

#define BOOST_VERSION_NUMBER(major, minor, patch) \
    ((((major) % 100) * 10000000) + (((minor) % 100) * 100000) + ((patch) % 100000))
#define BOOST_VERSION_NUMBER_MAX BOOST_VERSION_NUMBER(99, 99, 99999)
#define BOOST_VERSION_NUMBER_ZERO BOOST_VERSION_NUMBER(0, 0, 0)
#define BOOST_VERSION_NUMBER_MIN BOOST_VERSION_NUMBER(0, 0, 1)
#define BOOST_VERSION_NUMBER_AVAILABLE BOOST_VERSION_NUMBER_MIN
#define BOOST_VERSION_NUMBER_NOT_AVAILABLE BOOST_VERSION_NUMBER_ZERO

/* decimal decoders used by alpaka's BoostPredef.hpp / ApiCudaRt.hpp */
#define BOOST_PREDEF_MAKE_10_VVRRP(V) BOOST_VERSION_NUMBER(((V) / 1000) % 100, ((V) / 10) % 100, (V) % 10)
#define BOOST_PREDEF_MAKE_YYYYMMDD(V) BOOST_VERSION_NUMBER((((V) / 10000) % 10000) % 1900, ((V) / 100) % 100, (V) % 100)

#endif
