// timer.hpp -- device timing for the SpMV registry.
//
// Keeps the interface of reference/include/timer.hpp:8-67 (Meyers singleton; total_* /
// kernel_* start/stop; *_cost() in microseconds as int64) so main.cu's timing loop keeps its
// shape, but the clock is a pair of cudaEvents recorded on the SpMV stream instead of
// std::chrono on the host: the reference stops its "kernel" timer without a device sync in
// four of nine kinds, so its numbers are launch latencies.  *_cost()
// synchronises on the stop event, so it is safe to call right after SpMV() returns.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#define KERNEL_TIMER

// The stream every kind enqueues on and every Timer event is recorded on (the reference uses the
// legacy default stream everywhere; so does this, unless the caller sets another one).  One
// setting for both, so an event can never bracket work that runs on another stream.
struct SpmvStream {
    static cudaStream_t &get() {
        static cudaStream_t s = nullptr;
        return s;
    }
    static void set(cudaStream_t s) { get() = s; }
};

class Timer {
public:
    static Timer &get_instance() {
        static Timer timer;
        return timer;
    }

    static void set_stream(cudaStream_t s) { SpmvStream::set(s); }

    static void total_start() { get_instance().record(0); }
    static void total_stop() { get_instance().record(1); }
    static void kernel_start() {
#ifdef KERNEL_TIMER
        get_instance().record(2);
#endif
    }
    static void kernel_stop() {
#ifdef KERNEL_TIMER
        get_instance().record(3);
#endif
    }

    // microseconds, fractional
    static double total_cost_us() { return get_instance().elapsed_us(0, 1); }
    static double kernel_cost_us() {
#ifdef KERNEL_TIMER
        return get_instance().elapsed_us(2, 3);
#else
        return 0.0;
#endif
    }
    // the reference's integer-microsecond accessors
    static int64_t total_cost() { return (int64_t)(total_cost_us() + 0.5); }
    static int64_t kernel_cost() { return (int64_t)(kernel_cost_us() + 0.5); }

    Timer(const Timer &single) = delete;
    const Timer &operator=(const Timer &single) = delete;

private:
    Timer() {
        for (auto &e : ev_) cudaEventCreate(&e);
    }
    ~Timer() {
        for (auto &e : ev_) cudaEventDestroy(e);
    }
    void record(int i) {
        cudaEventRecord(ev_[i], SpmvStream::get());
        recorded_[i] = true;
    }
    double elapsed_us(int a, int b) {
        if (!recorded_[a] || !recorded_[b]) return 0.0;
        float ms = 0.f;
        cudaEventSynchronize(ev_[b]);
        if (cudaEventElapsedTime(&ms, ev_[a], ev_[b]) != cudaSuccess) return 0.0;
        return (double)ms * 1e3;
    }

    cudaEvent_t ev_[4]{};
    bool recorded_[4]{};
};
