// kernel/ntt.h -- the reference's kernel API (include/kernel/ntt.h:27-45), same names, argument order and
// meaning, served by libagxntt.so.  A driver that includes "kernel/ntt.h" and calls
//     ntt_input_kernel(...); fwd_ntt_kernel<0>(q); ntt_output_kernel(...); q.wait();
// (src/main.cpp:60-74) builds against this header unchanged.
//
// Differences, all deliberate: the transform size is taken at run time from the twiddle buffer's length instead of
// the compile-time FPGA_NTT_SIZE (ntt.h:7-23; the macro is still honoured as a check when defined), and any
// power of two from 4 to 32768 is accepted, not only {32, 1024, 8192, 16384, 32768}.
#ifndef AGX_KERNEL_NTT_H
#define AGX_KERNEL_NTT_H

#include <CL/sycl.hpp>
#include <sycl/ext/intel/fpga_extensions.hpp>

#ifndef FPGA_NTT_SIZE
#define FPGA_NTT_SIZE 16384
#endif

using namespace cl::sycl;

template <size_t idx> class FWD_NTT;

// compute-unit `id` of the reference (only 0 is instantiated there, ntt.cpp:648)
template <size_t id> void fwd_ntt_kernel(sycl::queue& q);

void ntt_input_kernel(buffer<uint64_t, 1>& inData_buf, buffer<uint64_t, 1>& inData2_buf,
                      buffer<uint64_t, 1>& modulus_buf, buffer<uint64_t, 1>& twiddleFactors_buf,
                      buffer<uint64_t, 1>& barrettTwiddleFactors_buf, unsigned int numFrames, sycl::queue& q);

void ntt_output_kernel(buffer<uint64_t, 1>& outData_buf, int numFrames, sycl::queue& q);

void fwd_ntt(sycl::queue& q);   // defined (undeclared) in the reference, ntt.cpp:643-645

#endif  // AGX_KERNEL_NTT_H
