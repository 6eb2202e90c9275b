// CL/sycl.hpp -- host-side SYCL look-alike that lets a driver written against the reference's kernel API
// (include/kernel/ntt.h:32-45; caller src/main.cpp:14-89) build with plain g++ and run on the B200 library.
//
// Only the surface main.cpp touches is provided: buffer<T,1>(size), host_accessor(buf, write_only|read_only),
// queue(selector, async_handler), queue::wait().  A `queue` owns one agx_ctx (include/agxntt.h); buffers are
// page-locked host memory that the reference-named entry points hand to agx_ref_*.  As in SYCL, constructing a
// host_accessor on a buffer with work in flight synchronises first, so main.cpp:80 is safe even without q.wait().
// SYCL errors surface as C++ exceptions (the reference's entry points return void): sycl::exception here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "agxntt.h"

namespace sycl {

class exception : public std::runtime_error {
    int code_;
public:
    exception(int code, const std::string& what)
        : std::runtime_error(what + ": " + agx_error_string(code)), code_(code) {}
    int code() const { return code_; }
};

struct exception_list {};
using async_handler = std::function<void(exception_list)>;
struct device_selector { virtual ~device_selector() = default; virtual int device() const { return 0; } };

class queue {
    struct State {
        agx_ctx* ctx = nullptr;
        // the round being assembled by the reference-named entry points (kernel/ntt.h): sizes seen so far, so that
        // whichever of ntt_input_kernel / ntt_output_kernel comes second can check the output buffer against numFrames x N
        size_t round_n = 0, round_out_words = 0;
        long long round_out_frames = -1;
        ~State() { if (ctx) agx_destroy(ctx); }
    };
    std::shared_ptr<State> st_;
    void init(int device) {
        st_ = std::make_shared<State>();
        const int rc = agx_create(&st_->ctx, nullptr, device);
        if (rc) throw exception(rc, "sycl::queue: no usable CUDA device");
    }
public:
    queue() { init(0); }
    explicit queue(const device_selector& s) { init(s.device()); }
    queue(const device_selector& s, const async_handler&) { init(s.device()); }
    agx_ctx* native() const { return st_->ctx; }
    // bookkeeping for ntt_shim.cpp; throws when the drain's buffer cannot hold numFrames x N words
    void note_input(size_t n) { st_->round_n = n; check_round(); }
    void note_output(size_t words, long long frames) { st_->round_out_words = words; st_->round_out_frames = frames; check_round(); }
    void end_round() { st_->round_n = 0; st_->round_out_frames = -1; }
private:
    void check_round() {
        if (st_->round_n && st_->round_out_frames >= 0) {
            const bool bad = (size_t)st_->round_out_frames * st_->round_n > st_->round_out_words;
            if (bad) { end_round(); throw exception(AGX_E_INVALID, "ntt_output_kernel: outData_buf is smaller than numFrames x N"); }
        }
    }
public:
    void wait() {
        end_round();
        const int rc = agx_wait(st_->ctx);
        if (rc) throw exception(rc, "sycl::queue::wait");
    }
    void wait_and_throw() { wait(); }
};

namespace detail {
// Buffer storage is page-locked host memory (agx_host_alloc) so that the loader and drain of the reference-shaped
// pipeline DMA straight from / into it and a whole round stays asynchronous until queue::wait(); if pinning fails
// (no device yet, limit reached) it falls back to ordinary memory, which the library stages through pinned chunks.
template <typename T> struct buffer_state {
    T* data = nullptr;
    size_t count = 0;
    bool pinned = false;
    std::unique_ptr<queue> pending;   // queue with work that reads/writes this buffer
    explicit buffer_state(size_t n) : count(n) {
        void* p = nullptr;
        if (n && agx_host_alloc(&p, n * sizeof(T)) == AGX_OK && p) { data = static_cast<T*>(p); pinned = true; }
        else data = static_cast<T*>(::operator new(n ? n * sizeof(T) : 1));
        std::memset(static_cast<void*>(data), 0, n * sizeof(T));
    }
    buffer_state(const buffer_state&) = delete;
    buffer_state& operator=(const buffer_state&) = delete;
    ~buffer_state() {
        sync_noexcept();
        if (pinned) agx_host_free(data); else ::operator delete(data);
    }
    void sync() { if (pending) { auto q = std::move(pending); q->wait(); } }
    void sync_noexcept() { try { sync(); } catch (...) {} }
};
}  // namespace detail

template <typename T, int D = 1> class buffer {
    static_assert(D == 1, "only 1-D buffers are needed by the NTT kernel API");
    std::shared_ptr<detail::buffer_state<T>> st_;
public:
    explicit buffer(size_t n) : st_(std::make_shared<detail::buffer_state<T>>(n)) {}
    size_t size() const { return st_->count; }
    size_t get_count() const { return size(); }
    T* host_data() { return st_->data; }
    bool is_pinned() const { return st_->pinned; }
    void mark_pending(const queue& q) { st_->pending.reset(new queue(q)); }
    void sync() { st_->sync(); }
};

struct write_only_t {};
struct read_only_t {};
struct read_write_t {};
inline constexpr write_only_t write_only{};
inline constexpr read_only_t read_only{};
inline constexpr read_write_t read_write{};

template <typename T, int D = 1> class host_accessor {
    T* p_;
    size_t n_;
public:
    template <typename Tag> host_accessor(buffer<T, D>& b, Tag) : p_(nullptr), n_(b.size()) {
        b.sync();                     // SYCL semantics: a host accessor waits for device work on the buffer
        p_ = b.host_data();
    }
    T& operator[](size_t i) const { return p_[i]; }
    size_t size() const { return n_; }
};
template <typename T, int D, typename Tag> host_accessor(buffer<T, D>&, Tag) -> host_accessor<T, D>;

}  // namespace sycl

namespace cl { namespace sycl = ::sycl; }
