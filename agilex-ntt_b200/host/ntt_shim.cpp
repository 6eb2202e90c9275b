// ntt_shim.cpp -- the reference's three entry points (src/kernel/ntt.cpp:87, :508, :610) over the C ABI.
#include "kernel/ntt.h"

namespace {
void check(int rc, const char* what) { if (rc) throw sycl::exception(rc, what); }
}

// Loader: hands the frame data, modulus and the two N-entry tables to the device side (ntt.cpp:508-607).
void ntt_input_kernel(buffer<uint64_t, 1>& inData_buf, buffer<uint64_t, 1>& inData2_buf,
                      buffer<uint64_t, 1>& modulus_buf, buffer<uint64_t, 1>& twiddleFactors_buf,
                      buffer<uint64_t, 1>& barrettTwiddleFactors_buf, unsigned int numFrames, sycl::queue& q) {
    const size_t N = twiddleFactors_buf.size();
    if (barrettTwiddleFactors_buf.size() != N || modulus_buf.size() < 1 ||
        inData_buf.size() < N * numFrames || inData2_buf.size() < N * numFrames)
        throw sycl::exception(AGX_E_INVALID, "ntt_input_kernel: buffer sizes do not describe numFrames x N");
    q.note_input(N);                  // with ntt_output_kernel's sizes: the drain's buffer must hold numFrames x N
    check(agx_ref_input(q.native(), (uint32_t)N, inData_buf.host_data(), inData2_buf.host_data(),
                        modulus_buf.host_data(), twiddleFactors_buf.host_data(),
                        barrettTwiddleFactors_buf.host_data(), numFrames),
          "ntt_input_kernel");
}

// Compute unit launch (ntt.cpp:86-506).
template <size_t id> void fwd_ntt_kernel(sycl::queue& q) { check(agx_ref_fwd(q.native(), (uint32_t)id), "fwd_ntt_kernel"); }
template void fwd_ntt_kernel<0>(sycl::queue& q);

void fwd_ntt(sycl::queue& q) { fwd_ntt_kernel<0>(q); }

// Drain: results land in outData_buf, row-major [numFrames][N] (ntt.cpp:610-640); complete after q.wait() or when
// a host_accessor is taken on the buffer.
void ntt_output_kernel(buffer<uint64_t, 1>& outData_buf, int numFrames, sycl::queue& q) {
    if (numFrames < 0) throw sycl::exception(AGX_E_INVALID, "ntt_output_kernel: negative numFrames");
    q.note_output(outData_buf.size(), numFrames);   // agx_ref_output takes a bare pointer: the size check lives here
    check(agx_ref_output(q.native(), outData_buf.host_data(), numFrames), "ntt_output_kernel");
    outData_buf.mark_pending(q);
}
