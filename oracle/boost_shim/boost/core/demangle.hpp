/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/core/demangle.hpp> (alpaka uses boost::core::demangle for
 * accelerator / kernel names in log lines only), see ../predef/version_number.h. */
#ifndef PPS_BOOST_SHIM_DEMANGLE_HPP
#define PPS_BOOST_SHIM_DEMANGLE_HPP

#include <cstdlib>
#include <cxxabi.h>
#include <string>

namespace boost
{
    namespace core
    {
        inline std::string demangle(char const* name)
        {
            int status = 0;
            char* p = abi::__cxa_demangle(name, nullptr, nullptr, &status);
            std::string s = (status == 0 && p) ? p : name;
            std::free(p);
            return s;
        }
    } // namespace core
} // namespace boost

#endif
