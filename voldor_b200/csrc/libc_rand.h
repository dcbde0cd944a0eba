// Speculative use of the process-wide libc rand() stream.
//
// The reference's mean-shift picks its start sample with `rand() % N` inside a loop that may stop early
// (meanshift.cu:73-97); how many numbers it consumes is part of its observable behaviour (the same unseeded stream
// feeds every later call).  To evaluate all trials in ONE kernel launch the draws must be known up
// front, so the stream is snapshotted, the maximum number of draws is taken, and — if the loop would have stopped
// early — the stream is rewound and advanced by exactly the number of draws the reference would have made.
// Uses only the POSIX random(3) state API (initstate/setstate), on which glibc's rand() is built.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace vb {

class LibcRandSnapshot {
public:
    // One-time check that rewinding really replays rand(): true on glibc, where rand() is random() on the state array
    // that initstate/setstate expose.  A libc whose rand() keeps private state (e.g. musl) fails the check and the
    // caller falls back to drawing one number at a time.  Leaves the stream exactly where it was.
    static bool supported() {
        static const bool ok = [] {
            LibcRandSnapshot probe;
            if (!probe.take()) return false;
            const int first = rand();
            probe.rewind();
            const int again = rand();
            probe.rewind();
            return first == again;
        }();
        return ok;
    }

    // Capture the generator state.  Returns false when the state array has an unexpected shape, in which case the
    // caller must fall back to drawing one number at a time.
    bool take() {
        live_ = initstate(1u, scratch_, sizeof(scratch_));  // park the generator on a scratch array
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
    // Put the stream back to the captured point.
    void rewind() {
        initstate(1u, scratch_, sizeof(scratch_));  // leave the live array before overwriting it
        memcpy(live_, saved_, size_);
        setstate(live_);
    }

private:
    char* live_ = nullptr;
    size_t size_ = 0;
    alignas(8) char saved_[256];
    alignas(8) char scratch_[256];
};

}  // namespace vb
