// The start-sample stream of the mean-shift: `rand() % N` inside a loop that may stop early
// (reference meanshift.cu:73-97).  How many numbers the loop consumes is part of the reference's observable
// behaviour: the same unseeded, process-wide libc stream feeds every later call.
//
// Two sources behind one interface:
//   * LibcStream    — the process-wide libc rand() itself.  Used by execution context 0, i.e. by every caller of the
//                     reference ABI, so a host program that seeds or draws from rand() sees exactly the reference's
//                     consumption.
//   * PrivateStream — a private generator producing the SAME sequence as glibc's rand() after srand(seed) (the
//                     TYPE_3 additive-feedback generator of random_r.c: r[i] = r[i-31] + r[i-3], seeded by the
//                     16807 Lehmer recurrence, first 310 outputs discarded).  Used by the additional execution
//                     contexts: each of them behaves like its own reference PROCESS (the reference's 6-worker process
//                     pool, slam_py/voldor_slam.py:182-187, gives every worker its own libc stream), and concurrent
//                     windows never race on shared generator state.
//
// Speculation: to evaluate all start-sample trials in ONE kernel launch the draws must be known up front, so the
// stream is snapshotted, the maximum number of draws is taken, and — if the reference's loop would have stopped
// early — the stream is rewound and advanced by exactly the number of draws the reference would have made.
//
// Thread-safety of LibcStream: rewinding goes through the POSIX random(3) state API (initstate/setstate), which is
// process-global.  Calls of this library are serialised per context, but a FOREIGN thread calling rand()/random()
// between snapshot() and rewind() would interleave with the speculation exactly as it would interleave with the
// reference's own rand() calls; such programs have no defined start-sample sequence in the reference either.  The
// parking array is static, so a concurrent rand() never touches a dead stack frame.  Set VB_NO_RAND_SPECULATION=1 to
// draw one number at a time instead (one launch per trial, like the reference).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace vb {

struct RandStream {
    virtual ~RandStream() {}
    virtual int next() = 0;         // the next rand() value
    virtual void seed(unsigned s) = 0;
    virtual bool snapshot() = 0;    // false: speculation unavailable, draw one number at a time
    virtual void rewind() = 0;      // back to the last snapshot
};

class LibcStream : public RandStream {
public:
    int next() override { return rand(); }
    void seed(unsigned s) override { srand(s); }
    // One-time check that rewinding really replays rand(): true on glibc, where rand() is random() on the state array
    // that initstate/setstate expose.  A libc whose rand() keeps private state (e.g. musl) fails the check and the
    // caller falls back to drawing one number at a time.  Leaves the stream exactly where it was.
    static bool supported() {
        static const bool ok = [] {
            if (getenv("VB_NO_RAND_SPECULATION")) return false;
            LibcStream probe;
            if (!probe.take()) return false;
            const int first = rand();
            probe.rewind();
            const int again = rand();
            probe.rewind();
            return first == again;
        }();
        return ok;
    }
    bool snapshot() override { return supported() && take(); }
    void rewind() override {
        initstate(1u, parking(), kStateBytes);  // leave the live array before overwriting it
        memcpy(live_, saved_, size_);
        setstate(live_);
    }

private:
    static constexpr size_t kStateBytes = 256;
    static char* parking() {
        alignas(8) static char scratch[kStateBytes];
        return scratch;
    }
    // Capture the generator state.  Returns false when the state array has an unexpected shape.
    bool take() {
        live_ = initstate(1u, parking(), kStateBytes);  // park the generator on the scratch array
        if (!live_) return false;
        // word 0 of a random(3) state array: MAX_TYPES(5) * rear-pointer index + type
        int32_t head;
        memcpy(&head, live_, sizeof(head));
        static const size_t kBytes[5] = {8, 32, 64, 128, 256};
        const int type = head % 5;
        const bool ok = head >= 0 && type >= 0 && type < 5;
        size_ = ok ? kBytes[type] : 0;
        if (ok) memcpy(saved_, live_, size_);
        setstate(live_);  // resume exactly where the stream was
        return ok;
    }
    char* live_ = nullptr;
    size_t size_ = 0;
    alignas(8) char saved_[kStateBytes];
};

class PrivateStream : public RandStream {
public:
    PrivateStream() { seed(1u); }  // an unseeded process starts from srand(1)
    void seed(unsigned s) override {
        if (s == 0) s = 1;
        int32_t word = (int32_t)s;
        st_.r[0] = word;
        for (int i = 1; i < kDeg; i++) {
            // word = 16807 * word mod (2^31 - 1) without overflow (Schrage)
            const long hi = word / 127773, lo = word % 127773;
            long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            word = (int32_t)w;
            st_.r[i] = word;
        }
        st_.f = kSep, st_.b = 0;
        for (int i = 0; i < kDeg * 10; i++) (void)next();
    }
    int next() override {
        const uint32_t v = (uint32_t)st_.r[st_.f] + (uint32_t)st_.r[st_.b];
        st_.r[st_.f] = (int32_t)v;
        if (++st_.f >= kDeg) st_.f = 0;
        if (++st_.b >= kDeg) st_.b = 0;
        return (int)(v >> 1);
    }
    bool snapshot() override {
        saved_ = st_;
        return getenv("VB_NO_RAND_SPECULATION") == nullptr;
    }
    void rewind() override { st_ = saved_; }

private:
    static constexpr int kDeg = 31, kSep = 3;
    struct State {
        int32_t r[kDeg];
        int f, b;
    } st_, saved_;
};

}  // namespace vb
