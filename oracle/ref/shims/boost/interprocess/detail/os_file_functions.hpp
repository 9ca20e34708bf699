// Stand-in for <boost/interprocess/detail/os_file_functions.hpp>: only named in a _WIN32 branch
// (openvdb/io/Archive.cc:493-503); the namespaces must exist for the using-directives.
#pragma once
namespace boost { namespace interprocess { namespace detail {} namespace ipcdetail {} }}
