// rb_session_plan.h — host-side bookkeeping of a streaming session (plain C++, no CUDA): which mixer frames the next
// render can produce from the PCM pushed so far, which part of every stream's FIFO the block reads, and what is left
// for the next block.  Shared by rb_session.cu and by the CPU emulator under tests/emu/, so the whole block logic is
// exercised on the CPU against whole-stream renders.
//
// What it replaces: rodio pulls one sample at a time through MixerSource::next (src/mixer.rs:120-136) ->
// UniformSourceIterator::next (src/source/uniform.rs:78-97) -> SampleRateConverter::next
// (src/conversions/sample_rate.rs:131-201); every adapter keeps its state between calls.  A block render must carry
// exactly that state from block to block:
//   * resampler: the stream-absolute index of the next output frame (its numerator and left frame follow from it:
//     (o * from) mod to and floor(o * from / to)) and the input frames from that left frame on -- they stay in the
//     stream's FIFO; an output is only rendered once its right neighbour has arrived, or the stream has ended (the
//     last frame is then emitted raw, sample_rate.rs:187-199);
//   * biquad: x[n-1], x[n-2], y[n-1], y[n-2] (src/source/blt.rs:397-410) -- four floats per stream on the device; a filter
//     in FRONT of the conversion additionally has a position of its own in the input (Stream::fpos): it has consumed the
//     right neighbour of the last output already, so its two last outputs are the interpolation taps.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace session {

struct Stream {
    uint32_t from = 1, to = 1; // the source's rate pair, reduced (from <= to; 1:1 = already at the mixer's rate)
    uint64_t mix_start = 0;    // mixer frame the stream joins at (frame aligned, src/mixer.rs:175-183)
    uint64_t pushed = 0;       // input frames received so far
    uint64_t out_done = 0;     // output frames rendered so far (= stream-absolute index of the next one)
    uint64_t i0 = 0;           // stream-absolute index of the frame at the front of the FIFO (multiple of 4)
    bool eof = false;          // no more input will come
    bool held = false;         // declared when the session was created but not handed to the mixer yet (Mixer::add comes
                               // later): it takes pushes, renders nothing and holds nobody up until it is started
    int64_t follows = -1;      // held source that starts on the frame after source `follows` has played out -- the
                               // sources of a queue (src/queue.rs:128-192: the next sound begins where the current one ends)
    uint64_t fpos = 0;         // input frames [0, fpos) have been pulled by the converter (stream-absolute): up to the right
                               // neighbour of the last output -- a gain in front of the conversion is final for them
    bool front = false;        // the source's filter sits in front of the conversion, runs once per INPUT frame and has
                               // consumed exactly those frames: the FIFO must keep everything from fpos on
    uint64_t fill() const { return pushed - i0; }   // frames in the FIFO
};

// SampleRateConverter output length for L input frames on the reduced grid from:to, from != to
// (closed form of src/conversions/sample_rate.rs:157-199; rb_sample_rate_out_len in the C ABI is the same formula).
inline uint64_t out_total(uint64_t L, uint32_t from, uint32_t to) {
    if (L == 0) return 0;
    if (L == 1) return 1;
    const uint64_t ns = ((L - 1) * (uint64_t)to + from - 1) / from;   // first output whose left frame is L-1
    return ns + ((ns * (uint64_t)from < L * (uint64_t)to) ? 1 : 0);
}
// Outputs [0, n) whose right neighbour exists among L frames: n = ceil((L-1) * to / from).
inline uint64_t out_interp(uint64_t L, uint32_t from, uint32_t to) {
    return L < 2 ? 0 : ((L - 1) * (uint64_t)to + from - 1) / from;
}
// Output frames of stream `s` that exist so far (renderable): everything once it has ended, else the interpolated ones.
inline uint64_t out_ready(const Stream& s) {
    return s.eof ? out_total(s.pushed, s.from, s.to) : out_interp(s.pushed, s.from, s.to);
}
inline bool finished(const Stream& s) { return s.eof && s.out_done >= out_total(s.pushed, s.from, s.to); }

// How many mixer frames [T, T + n) can be rendered now (every unfinished stream must be able to supply its part).
// Returns 0 with *ended = true when no stream is left (MixerSource::next returns None, src/mixer.rs:129-135).
// A queued source can be scheduled as soon as the end of its predecessor is known, i.e. once that one has received all
// of its input: its first frame follows the predecessor's last.  (Chains resolve front to back.)
// A predecessor that has played out before the successor was queued (Player::append on a player whose queue has run dry:
// the sound starts at the current position) hands over at T, the frame rendered next -- never behind the timeline, where the
// successor could neither render its head nor let the session move on.
inline void resolve_queue(std::vector<Stream>& st, uint64_t T) {
    for (bool again = true; again;) {
        again = false;
        for (Stream& s : st) {
            if (!s.held || s.follows < 0) continue;
            const Stream& p = st[(size_t)s.follows];
            if (p.held || !p.eof) continue;
            s.mix_start = std::max(T, p.mix_start + out_total(p.pushed, p.from, p.to));
            s.held = false, again = true;
        }
    }
}

inline uint64_t renderable(const std::vector<Stream>& st, uint64_t T, uint64_t max_frames, bool* ended) {
    uint64_t n = max_frames;
    bool active = false, waiting = false;
    for (const Stream& s : st) {
        if (s.held) {             // may still be added: keeps the session alive, holds nobody up
            waiting = true;
            continue;
        }
        if (finished(s)) continue;
        active = true;
        const uint64_t upto = s.mix_start + out_ready(s);   // the stream can cover the timeline up to here
        n = std::min(n, upto > T ? upto - T : 0);
    }
    *ended = !active && !waiting;
    return active ? n : 0;        // nothing is playing: MixerSource::next() is None, the timeline stands still (mixer.rs:129-135)
}

// One stream's part of the block [T, T + n).
struct Part {
    uint64_t mix_start = 0;   // block-local timeline position of its first output
    uint64_t out_len = 0;     // outputs it contributes
    uint64_t o0 = 0;          // stream-absolute index of the first of them
    uint64_t n_int = 0;       // how many of them interpolate (the rest is the raw last frame)
    bool continues = false;   // more outputs will follow in later blocks
};
inline Part part_of(const Stream& s, uint64_t T, uint64_t n) {
    Part p;
    if (s.held) {
        p.continues = true;
        return p;
    }
    const uint32_t from = s.from, to = s.to;
    const uint64_t ready = out_ready(s);
    const uint64_t lo = std::max(T, s.mix_start + s.out_done), hi = std::min(T + n, s.mix_start + ready);
    if (lo >= hi) {
        p.continues = !finished(s);
        return p;
    }
    p.mix_start = lo - T, p.out_len = hi - lo, p.o0 = lo - s.mix_start;
    const uint64_t ni = out_interp(s.pushed, from, to);
    p.n_int = ni > p.o0 ? std::min(p.out_len, ni - p.o0) : 0;
    p.continues = !(s.eof && p.o0 + p.out_len >= out_total(s.pushed, from, to));
    return p;
}
// Mixer::add for a held source while the mixer has rendered T frames: it joins at the next frame (src/mixer.rs:175-183).
inline void start(Stream& s, uint64_t T) {
    if (s.held) s.held = false, s.mix_start = T;
}

// Skippable::skip / Player::skip_one / stop: the source's iterator returns None from now on.  The mixer's converter has pulled
// frames [0, fpos) of it (the right neighbour of the last rendered output included): those are its whole input now, the rest
// of the FIFO is forgotten.  A held source (or one that has rendered nothing yet: fpos = 0) ends empty.
inline void skip(Stream& s) {
    s.pushed = std::min(s.pushed, std::max(s.fpos, s.i0));
    s.eof = true;
    if (s.held) s.held = false, s.follows = -1, s.pushed = s.i0 = s.fpos = 0, s.out_done = 0, s.mix_start = 0;
}

// After the block: advance the stream and tell how many FIFO frames (from the front) are dead.
inline uint64_t advance(Stream& s, const Part& p) {
    if (p.out_len) {   // the converter (and a filter in front of it) stands behind the right neighbour of the block's last output
        const uint64_t ia = ((p.o0 + p.out_len - 1) * (uint64_t)s.from) / s.to;
        s.fpos = std::min(ia + 2, s.pushed);
    }
    s.out_done = p.out_len ? p.o0 + p.out_len : s.out_done;
    uint64_t left = std::min((s.out_done * (uint64_t)s.from) / s.to, s.pushed);   // left frame of the next output
    if (s.front) left = std::min(left, s.fpos);   // more than two input frames per output: the filter still needs the ones between
    const uint64_t keep_from = left & ~3ull;                                        // FIFO front stays 16-byte aligned
    const uint64_t drop = keep_from > s.i0 ? keep_from - s.i0 : 0;
    s.i0 += drop;
    return drop;
}

}  // namespace session
