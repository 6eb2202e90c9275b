// Host stand-in for <sycl/ext/intel/fpga_extensions.hpp> (oracle/_ref build only; see CL/sycl.hpp).
#pragma once
#include <CL/sycl.hpp>
#include <deque>

namespace sycl { namespace ext { namespace intel {

struct fpga_selector : ::sycl::device_selector {};
struct fpga_emulator_selector : ::sycl::device_selector {};

// One unbounded FIFO per (Id, T, depth) instantiation.  A blocking read of an empty pipe can never be
// satisfied in this single-threaded model, so it throws instead of hanging.
template <class Id, typename T, size_t depth = 0> class pipe {
    static std::deque<T>& fifo() { static std::deque<T> q; return q; }
public:
    static void write(const T& v) { fifo().push_back(v); }
    static T read() {
        auto& q = fifo();
        if (q.empty()) throw std::runtime_error("sycl pipe stand-in: blocking read on an empty pipe (deadlock)");
        T v = q.front();
        q.pop_front();
        return v;
    }
    static size_t pending() { return fifo().size(); }
    static void clear() { fifo().clear(); }
};

}}}  // namespace sycl::ext::intel
