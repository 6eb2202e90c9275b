// Host stand-in for oneAPI dev-utilities <dpc_common.hpp> (oracle/_ref build only).
#pragma once
#include <CL/sycl.hpp>
namespace dpc_common {
inline auto exception_handler = [](sycl::exception_list) {};
}
