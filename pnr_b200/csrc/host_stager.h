// host_stager.h -- copies between the device and ORDINARY (pageable) host buffers at the PCIe rate.
//
// The unmodified call site hands Frangi::frangi3d buffers from operator new[] (Advantra_plugin.cpp:2490-2494).
// cudaMemcpyAsync on pageable memory is staged by the driver through its own pinned buffer on ONE host thread
// (measured: 970 ms per 2048 x 2048 x 512 call instead of 276 ms with pinned buffers), and pinning the caller's
// 17 GB for the duration of one call costs more than it saves.  The stager keeps a ring of pinned slots and a few
// host threads: a device-to-host piece is DMA'd into a slot on the copy stream and a worker moves it to the
// caller's pages as soon as the slot's event has fired; a host-to-device piece is gathered into a slot by the
// workers and DMA'd from there.  Several slots are in flight, so the host-side memcpy runs on several cores beside
// the DMA.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace frangi {

class HostStager {
public:
    static constexpr size_t kSlotBytes = 32u << 20;

    ~HostStager() { shutdown(); }

    // lazily started on the first pageable call; device = the device whose events the workers wait on
    bool start(int device, int nslots, int nthreads)
    {
        if (!slots_.empty()) return true;
        device_ = device;
        if (cudaSetDevice(device) != cudaSuccess) return false;
        slots_.resize(nslots);
        for (auto& s : slots_) {
            if (cudaHostAlloc(&s.p, kSlotBytes, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) != cudaSuccess) {
                shutdown();
                return false;
            }
            free_.push_back(&s);
        }
        stop_ = false;
        for (int t = 0; t < nthreads; ++t) workers_.emplace_back([this] { work(); });
        return true;
    }

    bool pageable(const void* p) const
    {
        if (!p) return false;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
        return a.type == cudaMemoryTypeUnregistered;
    }

    // device -> pageable host, ordered on `stream`; returns once every piece is enqueued (not finished: see drain)
    cudaError_t d2h(void* dst, const void* src, size_t bytes, cudaStream_t stream)
    {
        for (size_t off = 0; off < bytes; off += kSlotBytes) {
            const size_t n = bytes - off < kSlotBytes ? bytes - off : kSlotBytes;
            Slot* s = acquire();
            cudaError_t e = cudaMemcpyAsync(s->p, (const char*)src + off, n, cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaEventRecord(s->ev, stream);
            if (e != cudaSuccess) { release(s); return e; }
            push({ s, (char*)dst + off, nullptr, n });
        }
        return cudaSuccess;
    }

    // pageable host -> device on `stream`: the workers gather groups of pieces into slots, this thread issues the DMAs
    // in order.  Returns when the last DMA is enqueued (the slots are recycled by the workers after their events).
    cudaError_t h2d(void* dst, const void* src, size_t bytes, cudaStream_t stream)
    {
        struct Piece { Slot* s; size_t off, n; };
        const size_t group = slots_.size() / 2 ? slots_.size() / 2 : 1;   // half of the ring gathers while the rest drains
        size_t off = 0;
        while (off < bytes) {
            std::vector<Piece> g;
            for (size_t k = 0; k < group && off < bytes; ++k) {
                const size_t n = bytes - off < kSlotBytes ? bytes - off : kSlotBytes;
                Slot* s = acquire();
                s->gathered = false;
                push({ s, nullptr, (const char*)src + off, n });
                g.push_back({ s, off, n });
                off += n;
            }
            for (const Piece& q : g) {
                wait_gathered(q.s);
                cudaError_t e = cudaMemcpyAsync((char*)dst + q.off, q.s->p, q.n, cudaMemcpyHostToDevice, stream);
                if (e == cudaSuccess) e = cudaEventRecord(q.s->ev, stream);
                if (e != cudaSuccess) return e;
                push({ q.s, nullptr, nullptr, 0 });        // a worker frees the slot once the DMA has read it
            }
        }
        return cudaSuccess;
    }

    // every queued piece has reached its destination
    void drain()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return jobs_.empty() && busy_ == 0; });
    }

    void shutdown()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_job_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
        for (auto& s : slots_) {
            if (s.ev) cudaEventDestroy(s.ev);
            if (s.p) cudaFreeHost(s.p);
        }
        slots_.clear();
        free_.clear();
        jobs_.clear();
    }

private:
    struct Slot {
        void* p = nullptr;
        cudaEvent_t ev = nullptr;
        bool gathered = false;
    };
    struct Job {
        Slot* slot;
        char* dst;         // d2h: the caller's memory (copy out of the slot after its event)
        const char* src;   // h2d gather: the caller's memory (copy into the slot)
        size_t n;          // 0 with both pointers null: only wait for the event, then free the slot
    };

    Slot* acquire()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_free_.wait(lk, [this] { return !free_.empty(); });
        Slot* s = free_.back();
        free_.pop_back();
        return s;
    }
    void release(Slot* s)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            free_.push_back(s);
        }
        cv_free_.notify_one();
    }
    void push(const Job& j)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(j);
        }
        cv_job_.notify_one();
    }
    void wait_gathered(Slot* s)
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_gather_.wait(lk, [s] { return s->gathered; });
    }

    void work()
    {
        cudaSetDevice(device_);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
                if (stop_ && jobs_.empty()) return;
                j = jobs_.front();
                jobs_.pop_front();
                ++busy_;
            }
            if (j.src) {                                   // gather for an upload
                std::memcpy(j.slot->p, j.src, j.n);
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    j.slot->gathered = true;
                    --busy_;
                }
                cv_gather_.notify_all();
                cv_done_.notify_all();
                continue;
            }
            cudaEventSynchronize(j.slot->ev);
            if (j.dst) std::memcpy(j.dst, j.slot->p, j.n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                free_.push_back(j.slot);
                --busy_;
            }
            cv_free_.notify_one();
            cv_done_.notify_all();
        }
    }

    int device_ = 0;
    std::vector<Slot> slots_;
    std::vector<Slot*> free_;
    std::deque<Job> jobs_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_free_, cv_gather_, cv_done_;
    int busy_ = 0;
    bool stop_ = false;
};

}  // namespace frangi
