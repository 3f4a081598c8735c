// Build shim (test infrastructure): boost::filesystem -> std::filesystem.
#ifndef WSR_SHIM_BOOST_FS_H
#define WSR_SHIM_BOOST_FS_H
#include <filesystem>
namespace boost { namespace filesystem = std::filesystem; }
#endif
