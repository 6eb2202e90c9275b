// sycl/ext/intel/fpga_extensions.hpp -- selector names main.cpp uses (src/main.cpp:16-20), mapped to CUDA device 0
// (or $AGX_DEVICE).  There are no pipes here: the loader/compute/drain kernels of the reference are one CUDA kernel.
#pragma once
#include <cstdlib>
#include <CL/sycl.hpp>

namespace sycl { namespace ext { namespace intel {

struct fpga_selector : ::sycl::device_selector {
    int device() const override { const char* e = std::getenv("AGX_DEVICE"); return e ? std::atoi(e) : 0; }
};
struct fpga_emulator_selector : fpga_selector {};

}}}  // namespace sycl::ext::intel
