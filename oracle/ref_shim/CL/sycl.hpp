// Host stand-in for <CL/sycl.hpp>: just enough of the SYCL 2020 surface for the reference's
// src/kernel/ntt.cpp (and a main.cpp-shaped driver) to compile with plain g++ and run on the CPU.
// TEST INFRASTRUCTURE (oracle/_ref build only).  Written from the SYCL spec's public interface; it is not
// part of the product and contains no reference code.
//
// Execution model: queue::submit runs the command group immediately on the calling thread; single_task runs
// the kernel functor inline.  Pipes (see sycl/ext/intel/fpga_extensions.hpp) are unbounded FIFOs, so the
// loader -> compute -> drain kernels of the reference can run one after another instead of concurrently.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace sycl {

namespace access {
enum class mode { read, write, read_write };
}

struct exception_list {};
using async_handler = std::function<void(exception_list)>;

class handler {
public:
    template <typename Name = void, typename F>
    void single_task(const F& f) { f(); }
};

struct device_selector {};

class queue {
public:
    queue() = default;
    template <typename Sel> explicit queue(const Sel&) {}
    template <typename Sel, typename H> queue(const Sel&, const H&) {}
    template <typename CG> void submit(CG cg) { handler h; cg(h); }
    void wait() {}
    void wait_and_throw() {}
};

template <typename T, int D = 1> class buffer;

template <typename T, access::mode M> class accessor {
    T* p_;
public:
    explicit accessor(T* p) : p_(p) {}
    template <access::mode MM = M, typename std::enable_if<MM == access::mode::read, int>::type = 0>
    const T& operator[](size_t i) const { return p_[i]; }
    template <access::mode MM = M, typename std::enable_if<MM != access::mode::read, int>::type = 0>
    T& operator[](size_t i) const { return p_[i]; }
};

template <typename T, int D> class buffer {
    std::shared_ptr<std::vector<T>> v_;
public:
    explicit buffer(size_t n) : v_(std::make_shared<std::vector<T>>(n)) {}
    size_t size() const { return v_->size(); }
    T* raw() const { return v_->data(); }
    template <access::mode M> accessor<T, M> get_access(handler&) { return accessor<T, M>(v_->data()); }
};

struct write_only_t {};
struct read_only_t {};
struct read_write_t {};
inline constexpr write_only_t write_only{};
inline constexpr read_only_t read_only{};
inline constexpr read_write_t read_write{};

template <typename T, int D = 1> class host_accessor {
    T* p_;
public:
    template <typename Tag> host_accessor(buffer<T, D>& b, Tag) : p_(b.raw()) {}
    T& operator[](size_t i) const { return p_[i]; }
};
template <typename T, int D, typename Tag> host_accessor(buffer<T, D>&, Tag) -> host_accessor<T, D>;

}  // namespace sycl

namespace cl { namespace sycl = ::sycl; }
